"""`dataset.processors` of the B200 build: CLIPProcessor runs on the device (clip_processor.py); every other name of
the reference package (dataset/processors/__init__.py: BaseProcessor, the video processors, load_processor) keeps
resolving to the reference's own files when its tree is on sys.path behind this one.

`dataset`, `dataset.tools` and `evaluation` are namespace packages in the reference (no __init__.py) and here, so the
two trees merge: modules that exist here (dataset.tools.mask_toolbox, evaluation.tools, evaluation.eval_ullava) are
found first, everything else (dataset.datasets, dataset.collators, dataset.builders, dataset.tools.functional_video,
...) is the reference's.  This package is a regular one in both trees, hence the explicit path extension."""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)

from dataset.processors.clip_processor import CLIPProcessor  # noqa: E402

__all__ = ["CLIPProcessor"]

try:  # reference modules, present only when /path/to/u-LLaVA is on sys.path (and their dependencies installed)
    from dataset.processors.base_processor import BaseProcessor  # noqa: E402,F401
    __all__.append("BaseProcessor")
except Exception:  # pragma: no cover - reference tree absent
    pass
try:
    from dataset.processors.video_processor import (VideoTrainProcessor, VideoEvalProcessor,  # noqa: E402,F401
                                                    GIFTrainProcessor)
    __all__ += ["VideoTrainProcessor", "VideoEvalProcessor", "GIFTrainProcessor"]
except Exception:  # pragma: no cover - decord / imageio / reference tree absent
    pass


def load_processor(name, cfg=None):
    """processor = load_processor("clip_image", cfg) (reference dataset/processors/__init__.py:15-22)."""
    if name == "clip_image":
        return CLIPProcessor.from_config(cfg)
    from utils.registry import registry
    return registry.get_processor_class(name).from_config(cfg)
