#!/usr/bin/env python
"""profiles/r02_work.json from the ncu --set full captures of the hot kernels (gpurun_out/prof_r02_<kernel>.ncu-rep): executed lane
instructions, L2 bytes, DRAM bytes and (fill) shared-memory atomics PER VIEW of the captured launch.  bench.py divides them by its live
phase times: the work-based bounds of the `roofline.work` object.   usage: python scripts/make_work_json.py <views of the captured launch>"""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
views = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
KERNELS = {  # kernel -> (bench phase, share of that phase's time this kernel accounts for)
    "k_dense_occlusion_walk": ("occlusion", 1.0),
    "k_fill_tiles": ("fill", 1.0),
    "k_dense_tree": ("tree", 1.0),
    "k_dense_frustum": ("frustum", 0.9),
    "k_setup_triangles": ("setup", 0.85),
}
WANT = {"lanes_per_inst": "smsp__thread_inst_executed_per_inst_executed.ratio", "lts_sectors": "lts__t_sectors.sum", "dram_r": "dram__bytes_read.sum",
        "dram_w": "dram__bytes_write.sum", "ms": "gpu__time_duration.sum", "warp_inst": "smsp__inst_executed.sum", "sm_mhz": "sm__cycles_elapsed.avg.per_second",
        "smem_atom_wavefronts": "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum", "smem_wavefronts": "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active", "warps_active_pct": "sm__warps_active.avg.pct_of_peak_sustained_active"}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "inst": 1.0, "": 1.0,
        "hz": 1e-6, "Khz": 1e-3, "Mhz": 1.0, "Ghz": 1e3, "cycle/nsecond": 1e3, "cycle/usecond": 1.0, "sector": 1.0, "%": 1.0}
out = {"views_per_launch": views, "sms": 148, "l2_bytes_per_clk": 6300.0,
       "l2_source": "B300_MICROARCH.md: LTS throughput cap ~6300 B/clk (path-independent), same L2 organisation on B200", "kernels": {}}
for k, (phase, share) in KERNELS.items():
    rep = os.path.join(ROOT, "gpurun_out", "prof_r02_%s.ncu-rep" % k)
    if k == "k_dense_tree" and os.path.exists(rep.replace(".ncu-rep", "_main.ncu-rep")):
        rep = rep.replace(".ncu-rep", "_main.ncu-rep")  # the capture of the walk itself, not of its seed kernel
    if not os.path.exists(rep):
        continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    got = {}
    for key, name in WANT.items():
        if name in hdr:
            i = hdr.index(name)
            try:
                got[key] = float(vals[i].replace(",", "")) * UNIT.get(units[i], 1.0)
            except ValueError:
                pass
    ent = {"phase": phase, "phase_share": share, "captured_ms": got.get("ms"),
           "thread_inst_per_view": got.get("warp_inst", 0.0) * got.get("lanes_per_inst", 0.0) / views, "warp_inst_per_view": got.get("warp_inst", 0.0) / views,
           "lanes_per_inst": got.get("lanes_per_inst"), "issue_active_pct": got.get("issue_active_pct"), "warps_active_pct": got.get("warps_active_pct"),
           "lts_bytes_per_view": got.get("lts_sectors", 0.0) * 32.0 / views, "dram_bytes_per_view": (got.get("dram_r", 0.0) + got.get("dram_w", 0.0)) / views}
    if k == "k_fill_tiles" and "smem_atom_wavefronts" in got:
        ent["smem_atomics_per_view"] = got["smem_atom_wavefronts"] / views  # shared-memory wavefronts of atomic instructions
        ent["smem_wavefronts_per_view"] = got.get("smem_wavefronts", 0.0) / views
    if "sm_mhz" in got:
        out["sm_mhz"] = got["sm_mhz"]
    out["kernels"][k] = ent
    subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), rep, os.path.join(ROOT, "profiles", "r02_%s.txt" % k)], capture_output=True)
json.dump(out, open(os.path.join(ROOT, "profiles", "r02_work.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
