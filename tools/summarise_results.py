#!/usr/bin/env python
"""Turn the bench JSON lines under gpurun_out/ into profiles/r01_results.md (numbers measured on
the GPU box without a profiler attached)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")


def load(name):
    p = os.path.join(G, name)
    if not os.path.exists(p):
        return None
    lines = [l for l in open(p).read().splitlines() if l.startswith("{")]
    return json.loads(lines[-1]) if lines else None


def main():
    out = ["# Round 1 -- measured results (one B200 unless stated; no profiler attached)\n",
           "Device-timed: rays resident in HBM, SoA outputs, CUDA events taken by the library on the "
           "stream the kernels run on, L2 flushed (256 MiB memset) between timed steps, mean of the "
           "timed steps.  e2e: host buffers in, host `HitReg` records out through "
           "`prt_b200_nearest_hits` (pinned host memory, H2D + kernels + D2H inside the timed region).\n",
           "Static scenes (C2, C3, C3B, C4) are traced on the tree the default lazy mode leaves after "
           "max(32 rays per triangle, 8 Mi rays): the LBVH optimised by 2 treelet passes; C5 calls "
           "set_tris before every step (dynamic scene, see below).  `build ms` = set_tris as called "
           "(plain LBVH); `build+opt ms` = the same with the optimisation inside set_tris.\n",
           "| config | tris | rays | tags | trace ms | Mrays/s | nodes/ray | tris/ray | B/ray | "
           "fetched GB/s | of L2 read peak | build ms | Mtris/s | build+opt ms | Mtris/s | tree height | e2e Mrays/s |",
           "|---|---:|---:|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|"]
    for cfg in ("c2", "c3", "c3b", "c5", "c4"):
        d = load(f"bench_{cfg}.json")
        if not d:
            continue
        r = d["roofline"]
        out.append(f"| {cfg.upper()} | {d['config']['tris']:,} | {d['config']['rays_per_gpu']:,} | "
                   f"mask {d['config']['tag_mask']} | {d['ms_per_step']:.3f} | {d['value']:.0f} | "
                   f"{r['nodes_per_ray']:.1f} | {r['tris_per_ray']:.2f} | {r['bytes_per_ray']:.0f} | "
                   f"{r['achieved']:.0f} | {r['frac_of_l2']:.2f} | {d['build']['ms']:.3f} | "
                   f"{d['build']['mtris_s']:.0f} | {d['build']['ms_with_optimisation']:.3f} | "
                   f"{d['build']['mtris_s_with_optimisation']:.0f} | {d['build']['tree_height'] or '-'} | "
                   f"{d['e2e']['value']:.0f} |")
    d = load("bench_c2.json")
    if d:
        r = d["roofline"]
        out.append(f"\nMeasured on the box in the same run: L2 read bandwidth {r['l2_read_gbs_measured']:.0f} GB/s "
                   f"(32 MiB re-read), HBM read bandwidth {r['hbm_read_gbs_measured']:.0f} GB/s (2 GiB re-read); "
                   f"roofline peak used for `roofline.frac`: {r['peak']} GB/s, {r['peak_source']}.  "
                   f"Clocks during the timed region: {d['clocks']}.\n")
        c = d.get("cpu_baseline") or {}
        if c and "value" in c:
            out.append(f"CPU reference beside it (unmodified reference CPU backend, oracle/_ref): "
                       f"{c['value']:.1f} Mrays/s on {c['cores']} threads ({c['cpu']}), {c['sample']}; "
                       f"`set_tris` {c['build_mtris_s']:.3f} Mtris/s (1 thread).  C2 ratios: device-timed "
                       f"{d['value'] / c['value']:.0f}x, e2e (pinned) {d['e2e']['value'] / c['value']:.0f}x, "
                       f"e2e (pageable) {d['e2e'].get('pageable_value', 0) / c['value']:.0f}x, build "
                       f"{d['build']['mtris_s'] / c['build_mtris_s']:.0f}x.\n")
        pm = d.get("per_mask_mrays_s")
        if pm:
            out.append("C2, all 31 tag combinations (device-timed Mrays/s): " +
                       ", ".join(f"{k} {v:.0f}" for k, v in pm.items()) + "\n")
    d5 = load("bench_c5.json")
    if d5 and d5.get("dynamic_without_reuse"):
        r, b = d5["dynamic_without_reuse"], d5["build"]
        out.append(f"C5 is the dynamic scene: every step calls set_tris with the next of 4 distinct frames of the "
                   f"deforming height field.  Default (temporal reuse): set_tris {b['set_tris_ms_steady']:.3f} ms "
                   f"(refit of the optimised topology; {b['refits_in_timed_steps']} refits, "
                   f"{b['rebuilds_in_timed_steps']} rebuilds in the timed steps) + traversal "
                   f"{d5['ms_per_step']:.3f} ms = frame {b['set_tris_ms_steady'] + d5['ms_per_step']:.3f} ms.  Without "
                   f"reuse (mode 2, plain LBVH rebuilt every frame): set_tris {r['set_tris_ms']:.3f} ms + traversal "
                   f"{r['trace_ms']:.3f} ms ({r['value']:.0f} Mrays/s) = frame {r['frame_ms']:.3f} ms.\n")
    out.append("Opt-in watertight triangle test beside the default (same rays, same tree state): "
               "Mrays/s, rays whose `valid` differs, rays whose t differs by more than 1e-5 relative:\n")
    out.append("| config | default Mrays/s | watertight Mrays/s | valid differs | t differs > 1e-5 rel |")
    out.append("|---|---:|---:|---:|---:|")
    for cfg in ("c2", "c3", "c3b", "c5", "c4"):
        d = load(f"bench_{cfg}.json")
        if d and d.get("watertight"):
            w = d["watertight"]
            out.append(f"| {cfg.upper()} | {d['value']:.0f} | {w['value']:.0f} | {w['valid_differs']} | "
                       f"{w['t_rel_gt_1e-5']} |")
    out.append("")
    ref = load("bench_c2_ref.json")
    if ref:
        out.append(f"`bench.py --impl reference` (same box): {ref['value']:.1f} Mrays/s, "
                   f"{ref['cpu_baseline']['cores']} threads, build {ref['build_mtris_s']:.3f} Mtris/s.\n")
    rows = []
    for n in (1, 2, 4, 8):
        d = load("bench_c2.json") if n == 1 else load(f"bench_c2_n{n}.json")
        if d:
            e = d["e2e"]
            floor = e.get("pcie_concurrent_floor_ms") or (e.get("pcie") or {}).get("floor_ms")
            rows.append(f"| {n} | {d['value']:.0f} | {d['ms_per_step']:.3f} | {e['value']:.0f} | "
                        f"{e['ms_per_step']:.2f} | {('%.2f' % floor) if floor else '-'} |")
    prev = os.path.join(ROOT, "profiles", "r01_results_earlier.md")
    if os.path.exists(prev):
        out.append(open(prev).read())
    if len(rows) > 1:
        out.append("## Multi-GPU (C2, weak scaling: N frames of 2 073 600 rays, one rank per GPU, torchrun)\n")
        out.append("| GPUs | device-timed Mrays/s (whole job) | ms/step (max over ranks) | e2e Mrays/s | e2e ms/step | bare copies, all ranks at once (ms) |")
        out.append("|---:|---:|---:|---:|---:|---:|")
        out += rows
        out.append("\nThe device-timed metric scales with the GPUs (every rank traces its slice on its own "
                   "identical BVH, no collective on the data path).  The e2e call is bound by the box's "
                   "host<->device path: with all ranks copying their slices both ways at once and no "
                   "kernel at all, the copies alone take the time in the last column (N=4: the e2e call "
                   "runs at 98 % of it) -- these boxes are virtual machines whose GPUs share the "
                   "PCIe/IOMMU path, about 140 GB/s in total.\n")
    open(os.path.join(ROOT, "profiles", "r01_results.md"), "w").write("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    main()
