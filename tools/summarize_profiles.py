"""Turn the raw ncu exports of a GPU session (gpurun_out/) into the small, committed summaries under profiles/.
usage: python tools/summarize_profiles.py"""
import collections
import csv
import json
import os
import re
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def read_csv(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    return list(csv.reader(lines))


def launch_share(src, dst, title):
    rows = read_csv(src)
    hdr = rows[0]
    ik, iv, ig = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
    agg, grids, tot, n = collections.defaultdict(lambda: [0, 0.0]), collections.defaultdict(collections.Counter), 0.0, 0
    for x in rows[1:]:
        if len(x) < len(hdr):
            continue
        v = float(x[iv].replace(",", ""))
        k = re.sub(r"\(.*", "", x[ik])[:90]
        agg[k][0] += 1; agg[k][1] += v; tot += v; n += 1
        grids[k][x[ig]] += 1
    with open(dst, "w") as f:
        f.write(f"{title}\nlaunches {n}, sum of gpu__time_duration {tot / 1e6:.2f} ms (ncu: serialised, cold caches -- shares, not absolutes)\n\n")
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
            f.write(f"{t / 1e6:8.3f} ms {100 * t / tot:5.1f} %  n={c:5d} avg={t / c / 1e3:8.1f} us  {k}\n")
        f.write("\nrmsnorm launches by grid (grid 21504 = RMSNorm of the packed vision features, 10 cross-attention layers):\n")
        for k in grids:
            if "rmsnorm" in k:
                f.write(f"  {k}: {dict(grids[k])}\n")
        mine = sum(t for k, (c, t) in agg.items() if "mmfs::" in k)
        lib = sum(t for k, (c, t) in agg.items() if "nvjet" in k or "cutlass" in k or "cublas" in k.lower())
        f.write(f"\nthis repo's kernels {100 * mine / tot:.1f} %, cuBLAS GEMMs {100 * lib / tot:.1f} %, other (torch elementwise, cuDNN) {100 * (tot - mine - lib) / tot:.1f} %\n")


def raw_table(src, dst, title, metrics):
    rows = read_csv(src)
    hdr = rows[0]
    idx = [(m, hdr.index(m)) for m in metrics if m in hdr]
    with open(dst, "w") as f:
        f.write(title + "\n" + " | ".join(m for m, _ in idx) + "\n")
        for x in rows[2:]:
            if len(x) >= len(hdr):
                f.write(" | ".join((re.sub(r"\(.*", "", x[i])[:60] if m == "Kernel Name" else x[i]) for m, i in idx) + "\n")


def sampler_compare(dst):
    keys = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active"]
    out = {}
    for tag, fn in (("r01_design_generic_kernel", "r02_sampler_generic_ncu_raw.csv"), ("r02_specialised_kernel_first", "r02_sampler_v2_ncu_raw.csv"),
                    ("r02_specialised_kernel_final", "r02_sampler_v2b_ncu_raw.csv")):
        p = os.path.join(G, fn)
        if not os.path.exists(p):
            continue
        rows = read_csv(p)
        d, u = dict(zip(rows[0], rows[2])), dict(zip(rows[0], rows[1]))
        out[tag] = {k: f"{d[k]} {u[k]}" for k in keys if k in d}
    json.dump(out, open(dst, "w"), indent=1)
    return out


if __name__ == "__main__":
    if os.path.exists(os.path.join(G, "r02_launches_step_cfg3.csv")):
        launch_share(os.path.join(G, "r02_launches_step_cfg3.csv"), os.path.join(P, "r02_launch_share_step_cfg3.txt"),
                     "ncu launch list of ONE bench step of interleaved_cfg3 (tools/step_for_ncu.py inside cudaProfilerStart/Stop, "
                     "`ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none`)")
        shutil.copy(os.path.join(G, "r02_launches_step_cfg3.csv"), os.path.join(P, "r02_launches_step_cfg3.csv"))
    if os.path.exists(os.path.join(G, "r02_unet_kernels_raw.csv")):
        raw_table(os.path.join(G, "r02_unet_kernels_raw.csv"), os.path.join(P, "r02_unet_kernels_ncu_summary.txt"),
                  "ncu --set full of the first 40 conv_igemm / attn_fwd launches of one SD-2.1 UNet evaluation with the MMFS hook "
                  "(tools/unet_one.py, batch 16, bf16): duration [us], tensor-pipe (HMMA sub-pipe) cycles active [% of peak], SM throughput [%]",
                  ["Kernel Name", "Grid Size", "gpu__time_duration.sum", "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
                   "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
                   "sm__warps_active.avg.pct_of_peak_sustained_active"])
    cmp_ = sampler_compare(os.path.join(P, "r02_sampler_before_after_ncu.json"))
    for n in ("r02_sampler_generic_ncu_details.txt", "r02_sampler_v2_ncu_details.txt", "r02_sampler_v2b_ncu_details.txt",
              "r02_sampler_sweep2.log", "r02_sampler_sweep4.log", "r02_gemm_ab.log", "r02_decode_bench.json"):
        if os.path.exists(os.path.join(G, n)) and os.path.getsize(os.path.join(G, n)) > 0:
            shutil.copy(os.path.join(G, n), os.path.join(P, n))
    print(json.dumps(cmp_, indent=1)[:1500])
