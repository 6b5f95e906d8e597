"""Per-source-line summary of an ncu report captured with --import-source on:

    python profiles/source_lines.py gpurun_out/<name>.ncu-rep [top]

share of warp-stall samples, share of executed warp instructions and the two largest stall reasons
of the hottest source lines (ncu -i ... --page source --csv --print-source cuda,sass; inlined code
is attributed to its own file)."""
import collections
import csv
import subprocess
import sys


def main(path, top=40):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    fname, hdr = "", None
    lines = []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
        elif r[0] == "Line No":
            hdr = r
        elif hdr and r[0].isdigit():
            d = dict(zip(hdr, r))
            lines.append((fname, int(r[0]), r[1], r, hdr))
    tot_s = tot_i = 0
    out = []
    reasons_all = collections.Counter()
    for fname, ln, src, r, hdr in lines:
        vals = {}
        for h, v in zip(hdr, r):
            try:
                vals[h] = float(v)
            except ValueError:
                pass
        samples = vals.get("# Samples", 0.0)
        inst = vals.get("Instructions Executed", 0.0)
        stall = {h[len("stall_"):]: v for h, v in vals.items() if h.startswith("stall_")}
        tot_s += samples
        tot_i += inst
        reasons_all.update(stall)
        out.append((samples, inst, fname, ln, src.strip(), stall))
    out.sort(key=lambda t: -t[0])
    tot_r = sum(reasons_all.values()) or 1.0
    print("# all samples by reason: " + ", ".join("%s %.1f%%" % (k, 100 * v / tot_r) for k, v in reasons_all.most_common(9)))
    for samples, inst, fname, ln, src, stall in out[:top]:
        top2 = sorted(stall.items(), key=lambda kv: -kv[1])[:2]
        print("%-28s %4d  samples %4.1f%%  inst %4.1f%%  %-34s | %s"
              % (fname, ln, 100 * samples / max(tot_s, 1), 100 * inst / max(tot_i, 1),
                 ", ".join("%s %d" % (k, v) for k, v in top2), src[:110]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
