"""Regenerates the numeric sections of profiles/README.md from the tracked bench files (no GPU needed):
   profiles/r02_bench_1gpu.json, profiles/r02_bench_all_configs.jsonl, profiles/r01h_bench_all_configs.jsonl.
   python tools/make_profiles_readme.py > /tmp/sections.md   (the prose around the tables is kept by hand)"""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")


def jl(name):
    return [json.loads(l) for l in open(os.path.join(P, name)) if l.strip().startswith("{")]


LABEL = {
    "512x512 Fill - Naive": "C1 512x512 Naive, left-right (64 frames/step)",
    "1920x1080 Fill - Polylines Sharp": "C2 1080p Polylines Sharp + blur (16 frames/step)",
    "1920x1080 Imperfect fill - Hybrid Edge": "C3 1080p Hybrid Edge + blur",
    "3840x2160 GPU Warp (Fast),": "C4 4K GPU Warp, red-cyan anaglyph, div 10 (8 frames/step)",
    "7680x3840 Fill - Polylines Sharp": "C5 7680x3840 Polylines Sharp + blur, stereo_balance 0.5 (2 frames/step, 5 tiles per row)",
    "1920x1080 GPU Warp (Fast),": "1080p GPU Warp (node default, scatter warp)",
    "1920x1080 No fill,": "1080p No fill",
    "1920x1080 No fill - Reverse projection": "1080p No fill - Reverse projection",
    "1920x1080 Fill - Naive,": "1080p Naive",
    "1920x1080 Fill - Naive interpolating": "1080p Naive interpolating",
    "1920x1080 Fill - Polylines Soft": "1080p Polylines Soft",
    "1920x1080 Fill - Post-fill": "1080p none_post",
    "1920x1080 Fill - Reverse projection with Post-fill": "1080p inverse_post",
    "1920x1080 Fill - Hybrid Edge with fill": "1080p hybrid_edge_plus",
    "1920x1080 GPU Warp (Fast) [mesh warp]": "1080p GPU Warp, mesh warp (`COMFYSTEREO_GPU_WARP=mesh`)",
    "3840x2160 GPU Warp (Fast) [mesh warp]": "4K GPU Warp, mesh warp, red-cyan anaglyph, div 10",
}


def label(w):
    best = None
    for k in LABEL:
        if w.startswith(k) and (best is None or len(k) > len(best)):
            best = k
    return best


def main():
    r1 = {label(d["config"]["workload"]): d for d in jl("r01h_bench_all_configs.jsonl")}
    print("## Device-resident throughput (frames/s, one GPU)\n")
    print("| config (BASELINE.json) | fps | Mpix/s | path roofline | round 1 fps | dominant kernel |")
    print("|---|---|---|---|---|---|")
    for d in jl("r02_bench_all_configs.jsonl"):
        k = label(d["config"]["workload"])
        rf = d["roofline"]
        old = f"{r1[k]['value']:,.0f}" if k in r1 else "—"
        print(f"| {LABEL[k]} | {d['value']:,.0f} | {d['mpix_per_s']:,.0f} | {100 * rf['path']['frac']:.1f} % | {old} | "
              f"{rf['kernel']} {100 * rf['share_of_step']:.0f} % |")
    b = jl("r02_bench_1gpu.json")[-1]
    rf = b["roofline"]
    print(f"\nThe CPU oracle port of the reference on the box's {b['cpu_baseline']['cores']} cores: "
          f"{b['cpu_baseline']['value']:.1f} fps on the headline workload ({b['cpu_baseline']['sample']})\n"
          "(the reference itself — numba + PIL + three `gc.collect()` per call — measured 0.65 fps in the survey container).\n")
    print(f"## Headline step (1080p Polylines Sharp + blur, 16 frames per step): {b['value']:,.0f} fps, {b['ms_per_step']:.3f} ms\n")
    print("| kernel | ms/step | share | algorithmic GB/s |")
    print("|---|---|---|---|")
    for k, v in rf["kernels"].items():
        print(f"| {k} | {v['ms_per_step']:.3f} | {100 * v['share']:.0f} % | {v['gbs']:,.0f} |")
    print(f"\n`k_polylines` also composes (float32 side-by-side tensor + mask, 48 B/px algorithmic): {rf['achieved']:.0f} GB/s = "
          f"{100 * rf['frac']:.1f} % of the measured HBM peak ({rf['ms_per_launch']:.3f} ms per 16-frame launch); ncu: "
          f"{rf['traffic'] / 1e6:.0f} MB of DRAM traffic per launch.")
    print(f"Whole step: 80 B/px x 33.2 Mpx / {b['ms_per_step']:.3f} ms = {rf['path']['achieved']:.0f} GB/s = "
          f"{100 * rf['path']['frac']:.1f} % of the path roofline (round 1: 8.8 %).\n")
    e = b["e2e"]
    print(f"End to end through `StereoImageNode.generate` with page-locked host tensors: {e['value']:.0f} fps; pageable inputs (what\n"
          f"ComfyUI passes) {e['pageable_inputs']['value']:.0f} fps; 96 frames / 11 GB of results (pageable outputs) "
          f"{e['large_batch']['value']:.0f} fps.\nThe host's streaming-copy bandwidth measured in the same run is {e['host_copy_gbs']:.0f} GB/s; "
          f"the call moves\n{e['host_bytes_per_step'] / 16e6:.0f} MB of host memory traffic per frame = {100 * e['host_frac']:.0f} % of it: "
          "the host is the wall.\n")
    lat = b["single_frame_latency"]
    c0, c1 = lat["config0_512x512_naive"], lat["config1_1080p_polylines_sharp_blur"]
    print(f"Single-frame latency (device, CUDA events, graph replay): C1 {c0['median_us']:.0f} us per synchronised call, "
          f"{c0['back_to_back_us']:.0f} us back to back; C2 {c1['median_us']:.0f} / {c1['back_to_back_us']:.0f} us "
          "(round 1: 52 / 383 us per call).")


if __name__ == "__main__":
    main()
