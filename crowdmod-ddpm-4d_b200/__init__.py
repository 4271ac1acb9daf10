"""crowdmod-ddpm-4d_b200 — B200-native (sm_100a) hot path of marcemq/crowdmod-ddpm-4D.

Only what the path needs: ``csrc/`` (hand-written CUDA kernels + the C ABI declared in
``include/crowdmod_b200.h``), ``_native`` (ctypes binding) and a host-side mirror of the
reference's Python interface for this path (``models.backbones.unet.UNet``,
``models.diffusion.forward.ForwardSampler``, ``models.diffusion.ddpm.{DDPM, DDPM_model}``,
``utils.myparser``), same names / argument meaning / error behaviour.

The directory name carries a hyphen (it is the name the task fixes), so import it through the
alias module ``crowdmod_ddpm_4d_b200`` at the repo root, or put this directory itself on
``sys.path`` to shadow the reference's ``models`` / ``utils`` packages.
"""
__version__ = "0.1.0"
