"""Per-instruction stall samples / executed counts of one kernel from an `ncu --import-source on` report, and their sums over
named address ranges (the "phase" tables of DESIGN.md section 3 / 8).

    ncu -i gpurun_out/r02_step_full.ncu-rep --page source --csv --kernel-name regex:preprocess > /tmp/pre_src.csv
    python scripts/ncu_source_regions.py /tmp/pre_src.csv profiles/r02_ncu_source_preprocess.txt [first:last:label ...]

Ranges are instruction INDICES of the listing the script writes (column 1), so a second run can cut the phases once the
landmarks (barriers, MUFU, MATCH, strong loads ...) have been read off the first one."""
import csv
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    hdr, seen, out = None, set(), []
    for r in rows:
        if r and r[0] == "Address":
            hdr = r
            isrc, ismp, iex = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
            continue
        if hdr is None or len(r) <= iex or r[0] in seen:
            continue
        try:
            s, e = int(r[ismp] or 0), int(r[iex] or 0)
        except ValueError:
            continue
        seen.add(r[0])  # a report holding several launches of the kernel repeats the listing: keep the first
        out.append((r[0][-5:], s, e, r[isrc].strip()))
    return out


def main():
    out = load(sys.argv[1])
    tot_s, tot_e = sum(o[1] for o in out), sum(o[2] for o in out)
    lines = [f"# {sys.argv[1]}: {len(out)} SASS instructions, {tot_s} stall samples, {tot_e} warp instructions executed",
             "# index  address  stall_samples  warp_instructions_executed  SASS"]
    for rng in sys.argv[3:]:
        a, b, label = rng.split(":", 2)
        a, b = int(a), int(b)
        s, e = sum(o[1] for o in out[a:b]), sum(o[2] for o in out[a:b])
        lines.insert(1, f"# phase [{a:5d}, {b:5d})  samples {100 * s / max(tot_s, 1):5.1f} %  warp instructions {100 * e / max(tot_e, 1):5.1f} %  {label}")
    for i, o in enumerate(out):
        lines.append(f"{i:5d}  {o[0]}  {o[1]:6d}  {o[2]:9d}  {o[3]}")
    open(sys.argv[2], "w").write("\n".join(lines) + "\n")
    print("\n".join(l for l in lines if l.startswith("#")))


if __name__ == "__main__":
    main()
