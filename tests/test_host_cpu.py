"""CPU-only checks: the C-ABI library loads and exports every symbol include/ullava_sm100.h declares, the
host-side mirror keeps the reference's module API / state_dict keys, and the host logic of the sharded eval
path (shard ranges, result packing, the single all-gather under gloo with world_size 2)."""
import os
import re
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import native
    hdr = open(os.path.join(ROOT, "include", "ullava_sm100.h")).read()
    declared = set(re.findall(r"ULLAVA_API\s+[\w\s\*]+?\b(ullava_\w+)\s*\(", hdr))
    assert len(declared) >= 25
    assert declared == set(native.exported_symbols()), declared ^ set(native.exported_symbols())
    lib = native.load_library()  # dlopen; raises if the .so is missing (no fallback)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.ullava_abi_version() == native.ABI_VERSION == int(re.search(r"#define ULLAVA_ABI_VERSION (\d+)", hdr).group(1))


def test_struct_layouts_match_header_sizes():
    """ctypes mirrors of the C structs: spot-check sizes against the natural C layout."""
    import ctypes as C
    import native
    assert C.sizeof(native.GemmArgs) == 8 * 9 + 4 * 9 + 4  # 9 pointers/int64, 9 int32, tail padding
    assert C.sizeof(native.AttnArgs) == 16 * 8 + 4 * 9 + 4
    assert C.sizeof(native.DecodeArgs) > C.sizeof(native.LlamaArgs)


def test_header_is_plain_c_and_ctypes_mirrors_match_field_for_field(tmp_path):
    """include/ullava_sm100.h compiles as C99 with gcc (no CUDA, no C++), and every ctypes.Structure in native.py has
    the size AND the per-field offsets the C compiler gives the struct it mirrors."""
    import ctypes as C
    import shutil
    import subprocess
    import native
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    pairs = {"ullava_gemm_args": native.GemmArgs, "ullava_attn_args": native.AttnArgs,
             "ullava_sam_decoder_args": native.SamDecoderArgs, "ullava_vit_args": native.VitArgs,
             "ullava_sam_encoder_args": native.SamEncoderArgs, "ullava_llama_args": native.LlamaArgs,
             "ullava_decode_args": native.DecodeArgs}
    hdr = open(os.path.join(ROOT, "include", "ullava_sm100.h")).read()
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "ullava_sm100.h"', 'int main(void) {']
    for cname, cls in pairs.items():
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), hdr, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        fields = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            for part in decl.split(","):
                fields.append(re.findall(r"(\w+)\s*$", part.strip())[0])
        assert fields == [f[0] for f in cls._fields_], (cname, fields, [f[0] for f in cls._fields_])
        lines.append(f'  printf("{cname} %zu", sizeof({cname}));')
        for f in fields:
            lines.append(f'  printf(" %zu", offsetof({cname}, {f}));')
        lines.append('  printf("\\n");')
    lines += ['  return 0;', '}']
    src = tmp_path / "abi.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)],
                   check=True, capture_output=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.strip().splitlines()
    for line in out:
        name, size, *offs = line.split()
        cls = pairs[name]
        assert C.sizeof(cls) == int(size), (name, C.sizeof(cls), size)
        assert [getattr(cls, f[0]).offset for f in cls._fields_] == [int(o) for o in offs], name


def test_dataset_and_evaluation_merge_with_the_reference_tree():
    """With u-llava_b200/ in front of the reference tree on sys.path, `dataset` / `evaluation` stay namespace packages:
    the modules that exist here win, everything else still resolves to the reference's files (what
    inference_ullava.py:16-17 and tasks/setup_task need).  Skipped where /root/reference does not exist."""
    import importlib.util
    import subprocess
    ref = "/root/reference"
    if not os.path.isdir(ref):
        pytest.skip("reference tree not present")
    code = (
        "import sys, importlib.util as u\n"
        f"sys.path[:0] = [{os.path.join(ROOT, 'u-llava_b200')!r}, {ref!r}]\n"
        "f = lambda n: u.find_spec(n).origin\n"
        "for n in ['dataset.tools.mask_toolbox', 'dataset.processors.clip_processor', 'evaluation.tools',\n"
        "          'evaluation.eval_ullava', 'models', 'dataset.tools.functional_video',\n"
        "          'dataset.collators.base_collator', 'dataset.processors.video_processor', 'dataset.datasets.res_dataset']:\n"
        "    print(n, f(n))\n")
    out = subprocess.run([sys.executable, "-c", code], check=True, capture_output=True, text=True).stdout
    where = dict(line.split(" ", 1) for line in out.strip().splitlines())
    ours = os.path.join(ROOT, "u-llava_b200")
    for n in ("dataset.tools.mask_toolbox", "dataset.processors.clip_processor", "evaluation.tools",
              "evaluation.eval_ullava", "models"):
        assert where[n].startswith(ours), (n, where[n])
    for n in ("dataset.tools.functional_video", "dataset.collators.base_collator", "dataset.processors.video_processor",
              "dataset.datasets.res_dataset"):
        assert where[n].startswith(ref), (n, where[n])


def test_reference_inference_script_imports_resolve_to_this_build():
    """The import block of the reference's inference_ullava.py:12-20, executed with u-llava_b200/ in front of the
    reference tree: model / processor / toolbox names come from this build, utils.* from the reference, and the
    reference's registry hands out this build's classes ('ullava', 'clip_image').  peft is not installed in this
    image and is stubbed.  Skipped where /root/reference does not exist."""
    import subprocess
    ref = "/root/reference"
    if not os.path.isdir(ref):
        pytest.skip("reference tree not present")
    ours = os.path.join(ROOT, "u-llava_b200")
    code = f"""
import sys, types
sys.path[:0] = [{ours!r}, {ref!r}]
sys.modules['peft'] = types.ModuleType('peft'); sys.modules['peft'].PeftModel = object
from utils.tools import load_image
from transformers import LlamaTokenizer
from dataset.processors.clip_processor import CLIPProcessor
from dataset.tools.mask_toolbox import SegToolBox, DetToolBox
from utils.conversation import default_conversation, SeparatorStyle
from models import UllavaForCausalLM, KeywordsStoppingCriteria, \\
    DEFAULT_IMG_END_TOKEN, DEFAULT_IMG_START_TOKEN, DEFAULT_IMG_TOKEN, DEFAULT_IMG_PATCH_TOKEN
from models.ullava import UllavaForCausalLM as U2
from evaluation.tools import intersectionAndUnionGPU, AverageMeter, Summary, dict_to_cuda, bbox_iou
from utils.registry import registry
import models, utils.tools
assert U2 is UllavaForCausalLM
print(models.__file__); print(utils.tools.__file__)
print(registry.get_processor_class('clip_image').__module__, CLIPProcessor.__module__)
print(registry.get_model_class('ullava') is UllavaForCausalLM)
"""
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = r.stdout.strip().splitlines()
    assert lines[0].startswith(ours) and lines[1].startswith(ref), lines
    assert lines[2].split() == ["dataset.processors.clip_processor"] * 2 and lines[3] == "True", lines


def test_resize_longest_side_helpers_match_reference():
    """models.segment_anything.utils.transforms.ResizeLongestSide: shapes, coords and boxes equal the reference's
    (loaded from its file when the tree is present, otherwise the formulas are checked directly)."""
    import importlib.util
    import numpy as np
    from models.segment_anything.utils.transforms import ResizeLongestSide
    t = ResizeLongestSide(1024)
    ref_t = None
    ref_file = "/root/reference/models/segment_anything/utils/transforms.py"
    if os.path.exists(ref_file):
        try:
            spec = importlib.util.spec_from_file_location("ref_transforms", ref_file)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            ref_t = mod.ResizeLongestSide(1024)
        except Exception:
            ref_t = None
    rng = np.random.default_rng(1)
    for _ in range(40):
        h, w = int(rng.integers(20, 2000)), int(rng.integers(20, 2000))
        s = 1024.0 / max(h, w)
        assert t.get_preprocess_shape(h, w, 1024) == (int(h * s + 0.5), int(w * s + 0.5))
        boxes = rng.uniform(0, min(h, w), (5, 4))
        got = t.apply_boxes(boxes, (h, w))
        if ref_t is not None:
            assert ref_t.get_preprocess_shape(h, w, 1024) == t.get_preprocess_shape(h, w, 1024)
            assert np.array_equal(got, ref_t.apply_boxes(boxes, (h, w)))
            assert torch.equal(t.apply_boxes_torch(torch.from_numpy(boxes), (h, w)),
                               ref_t.apply_boxes_torch(torch.from_numpy(boxes), (h, w)))
        nh, nw = t.get_preprocess_shape(h, w, 1024)
        assert np.allclose(got[:, 0], boxes[:, 0] * (nw / w)) and np.allclose(got[:, 1], boxes[:, 1] * (nh / h))


def test_det_toolbox_matches_reference_arithmetic():
    """DetToolBox (dataset/tools/mask_toolbox.py:31-90): pad / normalise / denormalise round trip and the reference's
    formulas on random boxes; mask2bbox on a hand-made mask."""
    import numpy as np
    from dataset.tools.mask_toolbox import DetToolBox
    d = DetToolBox()
    rng = np.random.default_rng(3)
    for _ in range(50):
        w, h = int(rng.integers(50, 900)), int(rng.integers(50, 900))
        x0, y0 = rng.uniform(0, w / 2), rng.uniform(0, h / 2)
        box = [x0, y0, x0 + rng.uniform(1, w / 2), y0 + rng.uniform(1, h / 2)]
        side = max(w, h)
        px, py = ((0, (w - h) / 2.0) if w > h else ((h - w) / 2.0, 0))
        assert d.get_pad_length(w, h) == (px, py)
        n = d.pad_normalize_xyxy(box, w, h)
        assert n == [(box[0] + px) / side, (box[1] + py) / side, (box[2] + px) / side, (box[3] + py) / side]
        back = d.denormalize_padded_xyxy(n, w, h)
        assert np.allclose(back, box, atol=1e-9)
    assert d.xywh2xyxy([3, 4, 10, 20]) == [3, 4, 13, 24]
    m = np.zeros((20, 30), np.uint8)
    m[5:9, 7:15] = 1
    assert d.mask2bbox(m) == [7.0, 5.0, 14.0, 8.0]


def test_no_cuda_means_loud_failure():
    import native
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(native.NativeError):
        native.Context(0)


def test_state_dict_keys_match_reference():
    """tests/golden/*.json hold the state_dict key/shape lists of the REAL reference modules."""
    from tests.util_models import build_tiny_core, build_tiny_full, load_golden
    m, _, _ = build_tiny_core(torch.float32, device="cpu")  # strict=True load inside
    _, meta = load_golden("tiny_core")
    assert set(m.state_dict().keys()) == set(meta["shapes"].keys())
    m, _, _ = build_tiny_full(torch.float32, device="cpu")
    _, meta = load_golden("tiny_full")
    sd = m.state_dict()
    assert set(sd.keys()) == set(meta["shapes"].keys())
    for k, shp in meta["shapes"].items():
        assert list(sd[k].shape) == list(shp), k
    for attr in ("llm", "seg_projector", "det_projector", "det_decoder", "visual_model"):
        assert hasattr(m, attr)
    for attr in ("model", "lm_head", "vision_encoder", "vision_projector", "generate", "encode_image"):
        assert hasattr(m.llm, attr)
    for attr in ("image_encoder", "prompt_encoder", "mask_decoder", "postprocess_masks", "preprocess"):
        assert hasattr(m.visual_model, attr)


def test_config_roundtrip(tmp_path):
    import models
    from tests import configs as C
    cfg = models.UllavaConfig(llm_config=dict(C.TINY_LLM), seg_token_idx=C.SEG_ID, loc_token_idx=C.LOC_ID)
    d = cfg.to_dict()
    assert d["model_type"] == "ullava" and d["llm_config"]["model_type"] == "ullava_core"
    assert d["llm_config"]["vision_config"]["hidden_size"] == 64
    cfg.save_pretrained(tmp_path)
    back = models.UllavaConfig.from_pretrained(tmp_path)
    assert back.seg_token_idx == C.SEG_ID and back.llm_config.vision_hidden_layer == -2
    assert back.llm_config.mm_token_ids["IMG_START"] == C.MM_IDS["IMG_START"]
    from transformers import AutoConfig
    assert isinstance(AutoConfig.from_pretrained(tmp_path), models.UllavaConfig)


def test_interleave_gate_up_layout():
    from models.engine import interleave_gate_up
    g = torch.arange(32 * 3, dtype=torch.float32).view(32, 3)
    u = -g
    w = interleave_gate_up(g, u)
    assert w.shape == (64, 3)
    assert torch.equal(w[:16], g[:16]) and torch.equal(w[16:32], u[:16])
    assert torch.equal(w[32:48], g[16:]) and torch.equal(w[48:], u[16:])


def test_shard_ranges_cover_batch():
    from dist_eval import shard_range
    for gb, ws in ((256, 8), (10, 4), (3, 8), (32, 1)):
        spans = [shard_range(gb, r, ws) for r in range(ws)]
        assert spans[0][0] == 0 and spans[-1][1] == gb
        for a, b in zip(spans, spans[1:]):
            assert a[1] == b[0]
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1


def test_pack_unpack_results_roundtrip():
    from dist_eval import pack_results, unpack_bits, unpack_results
    g = torch.Generator().manual_seed(0)
    B, T, h, w = 3, 9, 5, 13
    words = (h * w + 31) // 32
    ids = torch.randint(0, 2 ** 40, (B, T), generator=g)
    masks = [torch.rand((n, h, w), generator=g) > 0.5 for n in (1, 0, 2)]
    bits = []
    for m in masks:
        flat = torch.zeros((m.shape[0], words * 32), dtype=torch.bool)
        flat[:, : h * w] = m.reshape(m.shape[0], h * w)
        wv = (flat.view(m.shape[0], words, 32).long() << torch.arange(32)).sum(-1)
        bits.append(torch.where(wv >= 2 ** 31, wv - 2 ** 32, wv).to(torch.int32))
    payload = pack_results(ids, bits, 2, words)
    ids2, bits2 = unpack_results(payload, T, 2, words)
    assert torch.equal(ids, ids2)
    for m, b in zip(masks, bits2):
        assert torch.equal(unpack_bits(b, h, w), m)


def _gather_worker(rank, ws, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws))
    sys.path[:0] = [ROOT, os.path.join(ROOT, "u-llava_b200")]
    import torch.distributed as dist
    from dist_eval import all_gather_results, shard_range
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    gb = 5  # ragged: shards of 3 and 2
    lo, hi = shard_range(gb, rank, ws)
    payload = torch.arange(lo, hi, dtype=torch.int32)[:, None] * 100 + torch.arange(7, dtype=torch.int32)[None]
    sizes = [shard_range(gb, r, ws)[1] - shard_range(gb, r, ws)[0] for r in range(ws)]
    out = all_gather_results(payload, sizes)
    expect = torch.arange(gb, dtype=torch.int32)[:, None] * 100 + torch.arange(7, dtype=torch.int32)[None]
    ret[rank] = bool(torch.equal(out, expect))
    dist.destroy_process_group()


def test_all_gather_results_gloo_world2():
    ws = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + os.getpid() % 2000
    mp.spawn(_gather_worker, args=(ws, port, ret), nprocs=ws, join=True)
    assert ret[0] and ret[1]


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm: oracle port on the host cores, bounded sample) runs without a GPU and
    prints ONE JSON line with the driver's keys."""
    import json
    import subprocess
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "images/s" and j["higher_is_better"] is True
    assert j["metric"].startswith("images/sec") and j["value"] > 0 and j["n_gpus"] == 1 and j["steps"] == 1
    assert j["e2e"] == {"value": j["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = j["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == j["value"] and len(cb["sample"]) > 20
    assert j["config"]["workload"].startswith("c5 shard") and j["vs_baseline"] is None


def test_lora_adapters_are_folded_into_packed_weights():
    """`model.llm = PeftModel.from_pretrained(model.llm, ...)` (inference_ullava.py:42-43): the injected LoRA layers keep
    the BASE weight in `.weight`; LlamaStack must pack W + scaling * B @ A, re-pack when the adapter state changes, and
    refuse adapter kinds it does not fold (never drop them silently)."""
    import torch
    from transformers import LlamaConfig, LlamaModel
    from models.engine import LlamaStack, effective_weight
    from tests.util_models import inject_lora
    torch.manual_seed(0)
    lc = LlamaConfig(vocab_size=64, hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=2,
                     num_key_value_heads=2)
    m = LlamaModel(lc).eval().to(torch.bfloat16)
    head = torch.nn.Linear(64, 64, bias=False).to(torch.bfloat16)
    stack = LlamaStack(m, head)
    stack.ensure()
    base_qkv = [stack.tensors[6 * l + 1].clone() for l in range(2)]
    deltas = inject_lora(m)
    m.to(torch.bfloat16)
    stack.ensure()                                         # module identity changed -> re-pack
    for l in range(2):
        a = m.layers[l].self_attn
        want = torch.cat([(a.q_proj.weight.float() + deltas[f"model.layers.{l}.self_attn.q_proj.weight"].delta()).to(torch.bfloat16),
                          a.k_proj.weight, (a.v_proj.weight.float() +
                                            deltas[f"model.layers.{l}.self_attn.v_proj.weight"].delta()).to(torch.bfloat16)], 0)
        assert torch.equal(stack.tensors[6 * l + 1], want)
        assert not torch.equal(stack.tensors[6 * l + 1], base_qkv[l])
    # disable_adapter(): the base model runs
    for l in range(2):
        m.layers[l].self_attn.q_proj.disable_adapters = m.layers[l].self_attn.v_proj.disable_adapters = True
    stack.ensure()
    assert all(torch.equal(stack.tensors[6 * l + 1], base_qkv[l]) for l in range(2))
    # merged flag: the delta is already in W
    q = m.layers[0].self_attn.q_proj
    q.disable_adapters = False
    q.weight.data += q.delta().to(torch.bfloat16)
    q.merged = True
    assert torch.equal(effective_weight(q), q.weight)
    # peft >= 0.6 layout: the wrapper holds the nn.Linear as .base_layer
    class NewStyle(torch.nn.Module):
        def __init__(self, inner):
            super().__init__()
            self.base_layer = torch.nn.Linear(inner.in_features, inner.out_features, bias=False).to(torch.bfloat16)
            self.lora_A, self.lora_B, self.scaling, self.r = inner.lora_A, inner.lora_B, inner.scaling, inner.r
            self.active_adapters, self.merged, self.disable_adapters = ["default"], False, False
    v = m.layers[1].self_attn.v_proj
    ns = NewStyle(v)
    assert torch.equal(effective_weight(ns), (ns.base_layer.weight.float() + v.delta()).to(torch.bfloat16))
    # DoRA / embedding adapters are refused
    ns.lora_magnitude_vector = {"default": torch.ones(4)}
    with pytest.raises(NotImplementedError):
        effective_weight(ns)
    # in-place .data edits leave no trace in the signature: invalidate() is the documented hook
    k = m.layers[1].self_attn.k_proj
    stack.ensure()
    k.weight.data.mul_(2)
    stack.ensure()
    stale = stack.tensors[6 + 1][64:128].clone()
    stack.invalidate()
    stack.ensure()
    assert not torch.equal(stack.tensors[6 + 1][64:128], stale)
