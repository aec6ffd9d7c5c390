#!/usr/bin/env python
"""Decode steps and SAM ViT-H encoder on the two lanes of an SM partition: each alone on its lane, and both at once.
One JSON line per split (GPU box):  python tools/bench_overlap.py 72 88 104 120"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "u-llava_b200")]


def child(sms):
    import torch
    import bench
    import native
    dev = torch.device("cuda", 0)
    model = bench.build_model(dev, torch.bfloat16)
    B, new = 32, 64
    ids, img, sam = (t.to(dev) for t in bench.make_inputs(0, B, torch.bfloat16))
    llm = model.llm
    part = native.Partition.get(0, sms) if sms else None
    main = torch.cuda.current_stream()
    enc = model.visual_model.image_encoder

    power = {}

    def timed(fn, reps=3, tag=None):
        fn()
        torch.cuda.synchronize()
        sampler = None
        if tag and os.environ.get("POWER"):
            sampler = bench.ClockSampler(0)
            sampler.start()
            reps = max(reps, 8)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        if sampler:
            import statistics
            sm, pw = [], []
            sampler.proc.terminate()
            for ln in sampler.lines:
                f = [x.strip() for x in ln.split(",")]
                try:
                    sm.append(float(f[0])); pw.append(float(f[2]))
                except Exception:
                    pass
            if sm:
                power[tag] = {"sm_mhz_median": statistics.median(sm), "power_w_median": statistics.median(pw),
                              "power_w_max": max(pw), "samples": len(sm)}
        return e0.elapsed_time(e1) / reps

    def gen(lane):
        kw = dict(_decode_lane=(part.ctx[0], part.streams[0])) if lane else {}
        llm.generate(input_ids=ids, images=img, max_new_tokens=new, do_sample=False, eos_token_id=-1, **kw)

    def sam_lane():
        st = part.streams[1]
        st.wait_stream(main)
        with torch.cuda.stream(st):
            enc.native_ctx = part.ctx[1]
            model.get_visual_embs(sam)
            enc.native_ctx = None
        main.wait_stream(st)

    sizes, resizes = [(bench.IMG, bench.IMG)] * B, [(bench.SAM_IMG, bench.SAM_IMG)] * B
    out = {"sms_decode": sms}
    if not sms:
        out["generate_full_machine_ms"] = timed(lambda: gen(False), tag="generate_full")
        out["sam_full_machine_ms"] = timed(lambda: model.get_visual_embs(sam), tag="sam_full")
        model.overlap_sam = False
        out["evaluate_ms"] = timed(lambda: model.evaluate(sam, img, ids, sizes, resizes, max_new_tokens=new, temperature=0))
    else:
        out["lanes"] = part.sms
        out["generate_decode_on_lane_ms"] = timed(lambda: gen(True), tag="generate_lane")
        out["sam_on_lane_ms"] = timed(sam_lane, tag="sam_lane")
        model.overlap_sms_decode = sms
        out["evaluate_overlapped_ms"] = {}
        for blocks in [int(b) for b in os.environ.get("OVERLAP_BLOCKS", "32").split(",")]:
            model.overlap_sam_blocks = blocks
            out["evaluate_overlapped_ms"][blocks] = round(timed(
                lambda: model.evaluate(sam, img, ids, sizes, resizes, max_new_tokens=new, temperature=0),
                tag=f"evaluate_overlapped_{blocks}"), 1)
    if power:
        out["power"] = power
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--child":
        child(int(sys.argv[2]))
    else:
        for s in [0] + [int(x) for x in sys.argv[1:]]:
            subprocess.run([sys.executable, __file__, "--child", str(s)], check=False)
