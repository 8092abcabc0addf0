"""Per-source-line summary of one kernel from an ncu report captured with --import-source on (build with -lineinfo):
stall samples and executed instructions aggregated per CUDA source line, top lines first.
usage: python tools/ncu_lines.py report.ncu-rep kernel_regex [top]"""
import csv, io, subprocess, sys


def main(path, kern, top=25):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    agg, order, hdr, cur, fname = {}, [], None, None, ""
    for r in rows:
        if len(r) >= 2 and r[0] == "Function Name":
            fname = r[1]; continue
        if len(r) > 6 and r[0] == "Line No":
            hdr = r; continue
        if hdr is None or len(r) < len(hdr):
            continue
        if r[0] != "-" and r[0] != "":
            cur = (r[0], r[1])
            if cur not in agg:
                agg[cur] = [0, 0, {}]; order.append(cur)
            if r[2] == "-":      # summary row of a source line
                try:
                    agg[cur][0] += int(r[hdr.index("# Samples")]); agg[cur][1] += int(r[hdr.index("Instructions Executed")])
                    for ci, name in enumerate(hdr):
                        if name.startswith("stall_") and "Not Issued" not in name:
                            agg[cur][2][name[6:]] = agg[cur][2].get(name[6:], 0) + int(r[ci] or 0)
                except ValueError:
                    pass
    tot_s = sum(v[0] for v in agg.values()) or 1
    tot_i = sum(v[1] for v in agg.values()) or 1
    print(f"{fname}\ntotal stall samples {tot_s}, warp instructions {tot_i}\n")
    print("| line | samples % | instr % | top stall reasons | source |\n|---|---|---|---|---|")
    for k in sorted(agg, key=lambda k: -agg[k][0])[:top]:
        s, i, st = agg[k]
        why = ", ".join(f"{n} {100 * v / max(1, s):.0f}%" for n, v in sorted(st.items(), key=lambda kv: -kv[1])[:3] if v)
        print(f"| {k[0]} | {100 * s / tot_s:.1f} | {100 * i / tot_i:.1f} | {why} | `{k[1].strip()[:100]}` |")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 25)
