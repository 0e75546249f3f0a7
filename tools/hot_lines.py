"""Per-source-line instruction and stall-sample totals from an ncu report (page source, view cuda,sass).
usage: python tools/hot_lines.py report.ncu-rep [top N]"""
import csv, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.check_output(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], text=True,
                              stderr=subprocess.DEVNULL)
rows = list(csv.reader(out.splitlines()))
lines = {}
func = ""
path = ""
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        path = r[1].split("/")[-1]
    if len(r) >= 2 and r[0] == "Function Name":
        func = path + " " + r[1][:40]
    if len(r) >= 8 and r[0].isdigit() and r[7].replace(",", "").isdigit():
        key = (func, int(r[0]))
        ins = int(r[7].replace(",", ""))
        smp = int(r[6]) if r[6].isdigit() else 0
        src = r[1].strip()
        a = lines.setdefault(key, [0, 0, src])
        a[0] += ins
        a[1] += smp
tot_i = sum(v[0] for v in lines.values())
tot_s = sum(v[1] for v in lines.values())
print("total warp instructions %d, samples %d" % (tot_i, tot_s))
for (f, ln), v in sorted(lines.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%-22s %5d  inst %5.1f%%  samples %5.1f%%  %s" % (f.split(" ")[0][:22], ln, 100.0 * v[0] / max(tot_i, 1), 100.0 * v[1] / max(tot_s, 1), v[2][:110]))
