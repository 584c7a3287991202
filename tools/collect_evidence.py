"""Copies the outputs of tools/gpu_final.sh (gpurun_out/f_* and gpurun_out/r02_*) into profiles/ under their judged names
and prints the headline numbers (to be quoted in DESIGN.md / profiles/README.md)."""
import json
import os
import shutil
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(REPO, "gpurun_out"), os.path.join(REPO, "profiles")
MAP = {"f_bench.json": "r02_bench_n1.json", "f_bench_ref.json": "r02_bench_reference_arm.json",
       "f_bench_bypass.json": "r02_bench_bypass_hfrm.json", "f_bench_cfg5.json": "r02_bench_cfg5.json",
       "f_bench_fp32.json": "r02_bench_fp32_tc32.json", "f_bench_fp32_ffma.json": "r02_bench_fp32_ffma.json",
       "f_bench_wiu.json": "r02_bench_wavelet_in_unet.json", "parity_s50.json": "r02_parity_s50.json",
       "f_pytest.log": "r02_pytest_gpu.log"}
R02 = ["launches_unet_p64.csv", "launch_summary.txt", "dram_unet_p64.csv", "traffic.json", "spans_events.txt", "ncu_full_1.csv",
       "ncu_full_2.csv", "ncu_full_3.csv", "ncu_full_4.csv", "dwt_dram.csv", "bench_dwt.txt", "hfrm_launches.csv",
       "hfrm_launch_summary.txt", "bench_hfrm.txt", "sampler_timeline.txt", "latency_small.txt", "lib_sha16.txt", "src_sha16.txt",
       "smi.txt", "p1_spans_splitk.txt", "latency_small_nosplit.txt", "train_step.json", "train_step.txt"]


def last_json(path):
    return json.loads(open(path).read().strip().splitlines()[-1])


def main():
    for a, b in MAP.items():
        shutil.copy(os.path.join(G, a), os.path.join(P, b))
    for f in R02:
        shutil.copy(os.path.join(G, "r02_" + f), os.path.join(P, "r02_" + f))
    print("lib sha16:", open(os.path.join(P, "r02_lib_sha16.txt")).read().strip())
    for name in ("r02_bench_n1", "r02_bench_bypass_hfrm", "r02_bench_cfg5", "r02_bench_fp32_tc32", "r02_bench_fp32_ffma",
                 "r02_bench_wavelet_in_unet", "r02_bench_reference_arm"):
        b = last_json(os.path.join(P, name + ".json"))
        r = b.get("roofline") or {}
        print(f"{name}: value {b['value']:.3f} e2e {b['e2e']['value']:.3f} ms {b['ms_per_step']:.1f} clk {(b.get('clocks') or {}).get('sm_mhz')} "
              f"frac {r.get('frac')} exec {r.get('executed_frac')} share {r.get('share_of_step')} traffic {r.get('traffic')} "
              f"hfrm {(b.get('hfrm') or {}).get('ms_per_batch')}")
    b = last_json(os.path.join(P, "r02_bench_n1.json"))
    print("traffic_src", b["roofline"]["traffic_src"])
    print("dwt", b["roofline_dwt"]["achieved"], b["roofline_dwt"]["frac"])
    print("cpu", b["cpu_baseline"])
    print("gpu eager", {k: (v["value"] if isinstance(v, dict) else v) for k, v in b["gpu_eager_baseline"].items() if k not in ("what",)})
    for k in ("fp32", "fp32_ffma", "bf16"):
        r = b["parity"][k]
        print(k, {x: r[x] for x in ("latent_max_rel", "image_max_abs", "psnr_abs_diff_db", "tc_launches", "simt_launches", "pass")})
    print(open(os.path.join(P, "r02_pytest_gpu.log")).read().strip().splitlines()[-2:])


if __name__ == "__main__":
    main()
