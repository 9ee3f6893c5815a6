#!/usr/bin/env python
"""Turn the JSON lines of studies/c5_suspicion_sweep.py into the markdown table kept under profiles/."""
import json
import sys


def main(path):
    rows = [json.loads(l) for l in open(path) if l.strip().startswith("{")]
    if not rows:
        print("no results in", path)
        return
    c = rows[0]["config"]
    print(f"# BASELINE config C5 — N = {c['n_nodes']:,} nodes on {c['n_gpus']} GPU(s), {c['rounds']} rounds, churn {c['crash_ppm']} ppm/round "
          f"(rejoin U[10,50]), {c['topology']} views, exchange {c['exchange']}, flags {c['flags']}\n")
    print("| S | suspicion_max | detection latency mean / p50 / p99 / max (rounds) | detected entries | undetected | stale Dead (down / up) | "
          "false Dead | suspected (up) | refutations | false-positive rate | mismatches at end | wall s | node-rounds/s (wall) |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
    for r in rows:
        c, e, l = r["config"], r["entries"], r["detection_latency_rounds"]
        lat = "-" if not l["n"] else f"{l['mean']:.1f} / {l['p50']} / {l['p99']} / {l['max']}"
        print(f"| {c['S']} | {c.get('suspicion_max', 0)} | {lat} | {e['detected']:,} | {e['undetected']:,} | {e['stale_dead']:,} / {e['stale_dead_up']:,} | "
              f"{e['false_dead']:,} | {e['suspected']:,} | {r['counters']['refutes']:,} | {r['false_positive_rate']:.2e} | "
              f"{r['mismatch_series'][-1][1]:,} | {r['wall_s']:.2f} | {r['node_rounds_per_s_wall']:.3e} |")
    print("\nMismatch count over time (entries of live observers that disagree with the truth), sampled:")
    for r in rows:
        s = r["mismatch_series"]
        pick = s[:: max(1, len(s) // 10)]
        print(f"* S = {r['config']['S']}: " + ", ".join(f"r{a}: {b:,}" for a, b in pick))
    print("\nsetup:", rows[0].get("setup_s"))


if __name__ == "__main__":
    main(sys.argv[1])
