"""Seed search behind SEED_TINY_CORE / SEED_TINY_FULL of make_golden.py (build container, CPU, oracle only).

The tiny random-weight models decide most greedy tokens by a top-2 logit margin of a few hundredths -- less than twice
the bf16 logit tolerance (2 * 8e-2), so an exact-id assertion against the reference could legitimately fail and the
tests of round 1 gated it off.  This script looks for weight seeds (oracle/synth.py: value = f(key, shape, seed)) for
which EVERY compared greedy decision has a margin above the threshold:

  core: 2 prompts x 8 tokens (tests/golden/tiny_core) + the cut prompt of the right-padded batch test (6 tokens)
  full: 2 [SEG]/[LOC] prompts x 8 tokens (evaluate() test: 6, smoke(): 8)

    python tests/golden/find_seed.py core 0 16000     # ~0.1 s per seed; seed 14400 -> 0.163
    python tests/golden/find_seed.py full 0 400       # seed 332 -> 0.252
    python tests/golden/find_seed.py coherent 0 2500  # CLIP-image seed of the coherent-mask fixture (oracle/synth.py): 2105
The goldens themselves are then written by make_golden.py from the REAL reference with those seeds.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "u-llava_b200"))

from oracle import ullava_oracle as O  # noqa: E402
from oracle.synth import synth_normal, synth_state_dict  # noqa: E402
from tests import configs as C  # noqa: E402
from tests.util_models import core_cfg, load_golden  # noqa: E402

torch.set_grad_enabled(False)


COH_SIZES = [(336, 336), (300, 420)]
COH_RESIZES = [(1024, 1024), (731, 1024)]


def coherent(lo, hi):
    """CLIP-image seeds (the SAM image stays at COHERENT_SEED, so its embedding is computed once) for which all three
    masks of the tiny_full prompts are two-signed and only their interpolated boundary lies near the threshold: a 16-bit
    evaluation moves a logit by ~1 % of the largest |logit|; a sharp two-level mask of this size has ~3e-4 of its
    pixels within 3 % of it (the boundary), a mask with a whole class of near-zero pixels has 1e-2."""
    from oracle.synth import COHERENT_SEED, coherent_image, coherent_overrides
    meta = load_golden("tiny_full")[1]
    cfg = core_cfg()
    sd = coherent_overrides(synth_state_dict(meta["shapes"], meta["seed"]))
    ids = C.tiny_prompt(2, seg_loc=True)
    emb = O.sam_image_encoder(sd, "visual_model.", coherent_image(2, seed=COHERENT_SEED), C.TINY_SAM_ENCODER)
    scfg = dict(seg_token_idx=C.SEG_ID, loc_token_idx=C.LOC_ID)
    for seed in range(lo, hi):
        ref = O.core_forward(sd, cfg, ids, synth_normal("images", (2, 3, 28, 28), seed=seed), prefix="llm.")
        pm, _, _ = O.masks_from_hidden(sd, scfg, ids, ref["last_hidden"], emb, COH_SIZES, COH_RESIZES)
        stats = [((x > 0).float().mean().item(), (x.abs() < 0.03 * x.abs().max()).float().mean().item())
                 for m in pm for x in m]
        if all(0.03 < p < 0.97 and n < 6e-4 for p, n in stats):
            print(f"CLIP-image seed {seed}: (positive share, share within 3 % of max |logit|) = "
                  f"{[(round(p, 3), round(n, 5)) for p, n in stats]}", flush=True)


def main():
    if sys.argv[1] == "coherent":
        return coherent(int(sys.argv[2]), int(sys.argv[3]))
    which, lo, hi = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    thr = float(sys.argv[4]) if len(sys.argv) > 4 else 0.16
    cfg = core_cfg()
    images = synth_normal("images", (2, 3, 28, 28))
    best = []
    if which == "core":
        shapes = load_golden("tiny_core")[1]["shapes"]
        ids = C.tiny_prompt(2)
        short = ids[1, : ids.shape[1] - 3]
        for seed in range(lo, hi):
            sd = synth_state_dict(shapes, seed)
            _, _, m1 = O.greedy_generate(sd, cfg, ids, images, 8)
            _, _, m2 = O.greedy_generate(sd, cfg, short[None], images[1:2], 6)
            best.append((min(m1.min().item(), m2.min().item()), seed))
    else:
        shapes = {k: v for k, v in load_golden("tiny_full")[1]["shapes"].items() if k.startswith("llm.")}
        ids = C.tiny_prompt(2, seg_loc=True)
        for seed in range(lo, hi):
            sd = synth_state_dict(shapes, seed)
            _, _, m = O.greedy_generate(sd, cfg, ids, images, 8, prefix="llm.")
            best.append((m.min().item(), seed))
    best.sort(reverse=True)
    for m, seed in best[:5]:
        print(f"seed {seed}: min margin {m:.4f}" + ("  <-- above threshold" if m > thr else ""))


if __name__ == "__main__":
    main()
