#!/usr/bin/env python
"""profiles/roofline_traffic.json from ncu CSVs (the `traffic` field of bench.py's roofline objects).

    python tools/traffic_from_ncu.py gpurun_out/r2_traffic_g0.csv [stream_gemm_raw.csv] > profiles/roofline_traffic.json

Input 1: `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct
-k regex:gemm_tcgen05 --csv python tools/gemm_traffic.py` -- two launches per shape of tools/gemm_traffic.SHAPES, the
second one is read.  `gemm_tensor` = DRAM bytes (read + write) per launch, averaged over the launch mix of one bench
step (32 LLaMA layers x {qkv, o, gate/up, down}, 32 SAM blocks x {qkv, proj, fc1, fc2}, 23 CLIP layers x 4 GEMMs; shapes
that were not captured are entered with their algorithmic operand + output bytes).
Input 2 (optional): an `ncu --page raw --csv` export holding gemm_stream_kernel launches; else the value measured in
round 1 is kept (the kernel is unchanged)."""
import csv
import json
import sys

NAMES = ["llama_qkv_b32", "llama_o_b32", "llama_gateup_b32", "llama_down_b32", "sam_qkv_b32", "sam_fc1_b32", "sam_fc2_b32",
         "vit_fc1_b32"]
# launches per bench step and algorithmic bytes (operands + output, 16-bit) of every large-M GEMM shape
MIX = {"llama_qkv_b32": (32, 19456, 12288, 4096, 12288), "llama_o_b32": (32, 19456, 4096, 4096, 4096),
       "llama_gateup_b32": (32, 19456, 22016, 4096, 11008), "llama_down_b32": (32, 19456, 4096, 11008, 4096),
       "sam_qkv_b32": (32, 131072, 3840, 1280, 3840), "sam_proj_b32": (32, 131072, 1280, 1280, 1280),
       "sam_fc1_b32": (32, 131072, 5120, 1280, 5120), "sam_fc2_b32": (32, 131072, 1280, 5120, 1280),
       "vit_qkv_b32": (23, 18464, 3072, 1024, 3072), "vit_o_b32": (23, 18464, 1024, 1024, 1024),
       "vit_fc1_b32": (23, 18464, 4096, 1024, 4096), "vit_fc2_b32": (23, 18464, 1024, 4096, 1024)}
STREAM_R1 = 117917952.0   # gemm_stream_kernel, DRAM bytes per launch, round-1 capture (profiles/r01_ncu_gemm_stream.md)


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
    hdr = rows[0]
    iid, imetric, ival = hdr.index("ID"), hdr.index("Metric Name"), hdr.index("Metric Value")
    per = {}
    for r in rows[1:]:
        per.setdefault(int(r[iid]), {})[r[imetric]] = float(r[ival].replace(",", ""))
    ids = sorted(per)
    measured = {}
    for k, nm in enumerate(NAMES):
        d = per[ids[2 * k + 1]]
        measured[nm] = d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"]
    total_bytes, total_launches, detail = 0.0, 0, {}
    for nm, (n, M, N, K, n_out) in MIX.items():
        algo = 2.0 * (M * K + N * K + M * n_out)
        b = measured.get(nm, algo)
        detail[nm] = {"launches_per_step": n, "dram_bytes_per_launch": b, "algorithmic_bytes": algo,
                      "measured": nm in measured, "over_read": round(b / algo, 2)}
        total_bytes += n * b
        total_launches += n
    out = {"gemm_tensor": total_bytes / total_launches, "gemm_stream": STREAM_R1,
           "source": f"tools/traffic_from_ncu.py {sys.argv[1]} (ncu dram__bytes_read.sum + dram__bytes_write.sum per launch; "
                     "launch-weighted mean over the large-M GEMMs of one bench step)", "per_shape": detail}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
