"""ncu launch list (--metrics gpu__time_duration.sum --csv) -> per-kernel totals and shares: python profiles/summarize_launches.py CSV OUT 'title'"""
import collections
import csv
import sys

src, out_path, title = sys.argv[1], sys.argv[2], sys.argv[3]
rows = [r for r in csv.reader(open(src)) if r and not r[0].startswith("==")]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt, order = collections.Counter(), collections.Counter(), []
for r in rows[1:]:
    if len(r) <= iv:
        continue
    ns = float(r[iv].replace(",", "")) * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(r[iu], 1)
    k = r[ik]
    if k not in tot:
        order.append(k)
    tot[k] += ns
    cnt[k] += 1
total = sum(tot.values())
lines = [f"# {title}", "# per-launch times are cold-cache and serialised: compare SHARES.", "# launches   total ns   share  kernel"]
lines += [f"{cnt[k]:4d} {tot[k]:14.0f} {100 * tot[k] / total:6.2f}%  {k[:110]}" for k in order]
step = {k: v for k, v in tot.items() if "sls_" in k and "kernel" in k and "peak" not in k or "best_reduce" in k}
if step:
    top = max(step, key=step.get)
    lines.append(f"# within the SLS steps ({' + '.join(sorted(x.split('(')[0].split('::')[-1] for x in step))}): {top.split('(')[0].split('::')[-1]} share = {100 * step[top] / sum(step.values()):.2f}%")
open(out_path, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
