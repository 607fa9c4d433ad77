"""Drop-in mirror of the reference's ``lib/model`` operator surface for the RoI /
proposal hot path.  Same dotted names, constructors and call
signatures; the work is done by libtlod_b200.so through ``tlod_b200``."""

# When the reference's lib/ directory follows this package's directory on sys.path (its
# _init_paths.py puts lib/ there), the reference's own sub-packages and modules that this mirror does
# not replace (model.faster_rcnn, model.rpn.rpn, model.utils.blob, ...) stay importable: the package
# path is extended with every other `model` directory on sys.path, this one first.
import pkgutil as _pkgutil

__path__ = _pkgutil.extend_path(__path__, __name__)
