# Lets this package shadow the reference's package of the same name on sys.path while the
# modules it does not replace still resolve to the reference's copies.
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
