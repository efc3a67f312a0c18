"""Dev tool: summarise a B200POST_CTA_TRACE dump (per-CTA start/end of one steady-state K2p launch)."""
import sys, collections
rows = [tuple(int(x) for x in l.split(",")) for l in open(sys.argv[1]) if l.strip()]
t0 = min(r[1] for r in rows)
dur = [(r[2] - r[1]) / 1e6 for r in rows]
end = [(r[2] - t0) / 1e6 for r in rows]
print(f"CTAs {len(rows)}  duration ms: min {min(dur):.2f} mean {sum(dur)/len(dur):.2f} max {max(dur):.2f};  launch span {max(end):.2f} ms; start skew {max((r[1]-t0)/1e6 for r in rows):.3f} ms")
per_sm = collections.defaultdict(list)
for r, d in zip(rows, dur): per_sm[r[3]].append(d)
sm_mean = sorted((sum(v) / len(v), k, len(v)) for k, v in per_sm.items())
print("SMs", len(per_sm), "CTAs/SM", collections.Counter(len(v) for v in per_sm.values()))
print("fastest SMs:", [(k, round(m, 2)) for m, k, n in sm_mean[:6]])
print("slowest SMs:", [(k, round(m, 2)) for m, k, n in sm_mean[-6:]])
within = [max(v) - min(v) for v in per_sm.values()]
print(f"within-SM CTA spread ms: mean {sum(within)/len(within):.2f} max {max(within):.2f}; across-SM mean spread {sm_mean[-1][0]-sm_mean[0][0]:.2f}")
import statistics
print("duration deciles:", [round(x, 2) for x in statistics.quantiles(dur, n=10)])
