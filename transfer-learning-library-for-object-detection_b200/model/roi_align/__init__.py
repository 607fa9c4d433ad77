
# When the reference's lib/ directory follows this package's directory on sys.path (its
# _init_paths.py puts lib/ there), the reference's own sub-packages and modules that this mirror does
# not replace (model.faster_rcnn, model.rpn.rpn, model.utils.blob, ...) stay importable: the package
# path is extended with every other `model/roi_align` directory on sys.path, this one first.
import pkgutil as _pkgutil

__path__ = _pkgutil.extend_path(__path__, __name__)
