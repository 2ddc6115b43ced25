"""One-screen summary of every kernel in an ncu report: time, instructions, issue rate, occupancy, DRAM bytes, top stalls."""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
def num(d, k):
    try: return float(d.get(k, "").replace(",", ""))
    except ValueError: return float("nan")
for r in rows[2:]:
    d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
    name = d["Kernel Name"].split("(")[0]
    stalls = []
    for k, v in d.items():
        if "pcsamp_warps_issue_stalled" in k and not k.endswith("not_issued"):
            try: stalls.append((float(v.replace(",", "")), k.split("stalled_")[1]))
            except ValueError: pass
    tot = sum(v for v, _ in stalls) or 1
    st = ", ".join(f"{k} {v / tot * 100:.0f}%" for v, k in sorted(stalls, reverse=True)[:5])
    print(f"{name:18s} {num(d, 'gpu__time_duration.sum'):7.3f} {u['gpu__time_duration.sum']:2s} inst {num(d, 'smsp__inst_executed.sum') / 1e6:8.1f}M issue {num(d, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):5.1f}% "
          f"warps {num(d, 'sm__warps_active.avg.pct_of_peak_sustained_active'):5.1f}% regs {num(d, 'launch__registers_per_thread'):.0f} "
          f"dram r/w {num(d, 'dram__bytes_read.sum'):.3g}/{num(d, 'dram__bytes_write.sum'):.3g} {u['dram__bytes_read.sum']}/{u['dram__bytes_write.sum']} L2hit {num(d, 'lts__t_sector_hit_rate.pct'):.0f}%")
    print(f"{'':18s} stalls: {st}")
