"""Drop-in `models` package for the u-LLaVA inference hot path on B200 (sm_100a).

Same importable names as /root/reference/models/__init__.py; put the directory that contains this
package (u-llava_b200/) in front of sys.path and `from models import UllavaForCausalLM` resolves here
(see INTEGRATION.md)."""
import os as _os
import sys as _sys

_pkg_root = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
if _pkg_root not in _sys.path:  # makes `import native` resolvable
    _sys.path.insert(0, _pkg_root)

from models.ullava_core import UllavaCoreConfig, UllavaCoreForCausalLM  # noqa: E402
from models.ullava import UllavaConfig, UllavaForCausalLM  # noqa: E402
from models.tools import (KeywordsStoppingCriteria, smart_resize_token_embedding,  # noqa: E402
                          smart_special_token_and_embedding_resize, multi_modal_resize_token_embedding)

DEFAULT_IMG_TOKEN = '<image>'
DEFAULT_IMG_PATCH_TOKEN = "<image_patch>"
DEFAULT_IMG_START_TOKEN = "<img_beg>"
DEFAULT_IMG_END_TOKEN = "</img_end>"
DEFAULT_VID_PATCH_TOKEN = "<video_patch>"
DEFAULT_VID_START_TOKEN = "<vid_beg>"
DEFAULT_VID_END_TOKEN = "</vid_end>"
DEFAULT_SEG_TOKEN = '[SEG]'
DEFAULT_LOC_TOKEN = '[LOC]'
DEFAULT_TAG_START = '[tag]'
DEFAULT_TAG_END = '[/tag]'
DEFAULT_BOS_TOKEN = '<s>'
DEFAULT_EOS_TOKEN = '</s>'
DEFAULT_UNK_TOKEN = '<unk>'
DEFAULT_PAD_TOKEN = '[PAD]'
IGNORE_INDEX = -100

__all__ = [
    "UllavaConfig", "UllavaForCausalLM", "UllavaCoreConfig", "UllavaCoreForCausalLM", "KeywordsStoppingCriteria",
    "smart_resize_token_embedding", "multi_modal_resize_token_embedding", "smart_special_token_and_embedding_resize",
    "DEFAULT_IMG_TOKEN", "DEFAULT_SEG_TOKEN", "DEFAULT_LOC_TOKEN", "DEFAULT_IMG_PATCH_TOKEN",
    "DEFAULT_IMG_START_TOKEN", "DEFAULT_IMG_END_TOKEN", "DEFAULT_VID_PATCH_TOKEN", "DEFAULT_VID_START_TOKEN",
    "DEFAULT_VID_END_TOKEN", "DEFAULT_BOS_TOKEN", "DEFAULT_EOS_TOKEN", "DEFAULT_UNK_TOKEN", "DEFAULT_PAD_TOKEN",
    "IGNORE_INDEX", "DEFAULT_TAG_START", "DEFAULT_TAG_END",
]
