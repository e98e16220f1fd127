"""Summarise an `ncu --page source --print-source cuda,sass --csv` export by source line:
    ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > src.csv
    python profiles/ncu_hot_lines.py src.csv [N]
"""
import csv
import sys


def main(path, top=30):
    rows = list(csv.reader(open(path)))
    cur_file, hdr, out = None, None, []
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            cur_file, hdr = r[1].split("/")[-1], None
            continue
        if len(r) > 3 and r[0] == "Line No":
            hdr = {h: i for i, h in enumerate(r)}
            continue
        if hdr is None or len(r) < len(hdr) or r[2] != "-":
            continue  # keep only the per-source-line summary rows (Address == "-")

        def f(name):
            try:
                return float(r[hdr[name]])
            except Exception:
                return 0.0

        out.append((cur_file, r[0], r[1].strip(), f("Instructions Executed"), f("# Samples"), f("stall_barrier"),
                    f("stall_long_sb"), f("stall_short_sb"), f("L1 Wavefronts Shared"), f("L1 Wavefronts Shared Ideal")))
    ti, ts = sum(a[3] for a in out), sum(a[4] for a in out)
    print(f"total warp-instructions {ti:.4g}, samples {ts:.0f}")
    print("== by instructions executed")
    for a in sorted(out, key=lambda a: -a[3])[:top]:
        print(f"{a[3] / ti * 100:5.1f}% inst {a[4] / ts * 100:5.1f}% samp | {a[0]}:{a[1]:>4} | {a[2][:100]}")
    print("== by stall samples")
    for a in sorted(out, key=lambda a: -a[4])[:top]:
        print(f"{a[4] / ts * 100:5.1f}% samp (bar {a[5]:.0f} long {a[6]:.0f} short {a[7]:.0f}) smem wf {a[8]:.3g}/{a[9]:.3g} | {a[0]}:{a[1]:>4} | {a[2][:90]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30)
