"""Import alias for the hyphenated package directory ``crowdmod-ddpm-4d_b200/``.

``import crowdmod_ddpm_4d_b200.models.backbones.unet`` resolves into that directory.
"""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "crowdmod-ddpm-4d_b200")]
with open(_os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), _f.name, "exec"))
del _f, _os
