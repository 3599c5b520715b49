"""ncu --page raw --csv export -> small JSON (per kernel: time, instructions, issue-slot utilisation, DRAM bytes, occupancy) that
bench.py embeds as `roofline.ncu_committed_capture`.  Usage: ncu_to_json.py raw.csv out.json [source note]"""
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
names = {"blend_forward_kernel": "blend_forward", "blend_backward_transposed_kernel": "blend_backward", "blend_backward_kernel": "blend_backward_butterfly",
         "preprocess_kernel": "preprocess", "onesweep_pass_kernel": "sort_pass", "backward_points_kernel": "backward_points",
         "sort_histogram_kernel": "sort_histogram", "tile_ranges_kernel": "tile_ranges"}
get = lambda r, k: float(r[hdr.index(k)].replace(",", "")) if k in hdr and r[hdr.index(k)] not in ("", "n/a") else None  # noqa: E731
out = {"source": sys.argv[3] if len(sys.argv) > 3 else sys.argv[1]}
for r in rows[2:]:
    kn = r[hdr.index("Kernel Name")]
    key = next((v for k, v in names.items() if k in kn), None)
    if key is None or key in out:
        continue
    rd, wr = get(r, "dram__bytes_read.sum"), get(r, "dram__bytes_write.sum")
    ur, uw = r[hdr.index("dram__bytes_read.sum")], None
    units = rows[1]
    scale = lambda name: {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}.get(units[hdr.index(name)], 1.0)  # noqa: E731
    out[key] = {
        "kernel": kn.strip(), "duration_us_under_ncu": get(r, "gpu__time_duration.sum"),
        "warp_instructions": get(r, "smsp__inst_executed.sum"),
        "issue_slots_active_pct": get(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "warps_active_pct_of_peak": get(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        "registers_per_thread": get(r, "launch__registers_per_thread"),
        "pipe_fma_pct": get(r, "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
        "pipe_alu_pct": get(r, "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
        "pipe_xu_pct": get(r, "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
        "pipe_lsu_pct": get(r, "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
        "dram_throughput_pct": get(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        "dram_bytes_per_launch": (rd * scale("dram__bytes_read.sum") + wr * scale("dram__bytes_write.sum")) if rd is not None and wr is not None else None,
        "l2_hit_rate_pct": get(r, "lts__t_sector_hit_rate.pct"),
    }
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(json.dumps(out, indent=1))
