"""Convergence-study measurements on top of the bulk simulator (BASELINE config C5: detection latency, false
positives and convergence against the suspicion timeout S). Pure host code over the arrays a handle exposes
(`get_array`, `counters`, `mismatches`): it works on any object with the Simulator's interface.

Terms, per view entry (observer i, member m) of a live observer:
  member down: `detected` = the entry is Dead and was set at or after the crash (latency = vlast - crash round);
               `stale_dead` = Dead since before the crash (an earlier death whose rejoin never reached i);
               otherwise `undetected` (still Alive or Suspect);
  member up:   `suspected` = Suspect (a transient false alarm, or a rejoin that is not known yet);
               `false_dead` = Dead and set after the member's last (re)start — declared dead while it was up;
               `stale_dead_up` = Dead since before the member's last rejoin (the Alive(inc+1) has not arrived).
The mismatch count of the library (`swim_sim_mismatches`) counts every entry that disagrees with the truth, i.e.
undetected + suspected + false_dead + stale_dead_up (a Dead entry of a down member agrees, stale or not)."""
import numpy as np

from . import _abi as A


def event_history(events, n_nodes, upto_round):
    """(last_crash, last_rejoin): the round of each node's last CRASH / REJOIN event with round <= upto_round (0 = none)."""
    last_crash = np.zeros(n_nodes, dtype=np.int64)
    last_rejoin = np.zeros(n_nodes, dtype=np.int64)
    if len(events):
        ev = events[events["round"] <= upto_round]
        order = np.argsort(ev["round"], kind="stable")
        ev = ev[order]
        c = ev[ev["kind"] == A.EV_CRASH]
        last_crash[c["node"]] = c["round"]        # later rounds overwrite earlier ones (sorted ascending)
        r = ev[ev["kind"] == A.EV_REJOIN]
        last_rejoin[r["node"]] = r["round"]
    return last_crash, last_rejoin


def view_report(sim, events, round_now, max_latency=256):
    """Classify every view entry of this rank's live observers (see the module docstring). Returns a dict of counts
    plus `latency_hist` (index = rounds from crash to the Dead mark, clipped to max_latency). events = None: the crash /
    rejoin rounds come from the handle's SWIM_ARR_LAST_CRASH / SWIM_ARR_LAST_REJOIN arrays (seeded device-side churn)."""
    cap = sim.cfg.view_cap
    alive = sim.get_array(A.ARR_ALIVE).astype(bool)                  # [N] truth, replicated on every rank
    n_total = alive.shape[0]
    nbr = sim.get_array(A.ARR_NBR).reshape(-1, cap)
    st = (sim.get_array(A.ARR_VST).reshape(-1, cap) & 3)
    vlast = sim.get_array(A.ARR_VLAST).reshape(-1, cap).astype(np.int64)
    first = getattr(sim, "first", 0)
    n_local = nbr.shape[0]
    if events is None:  # the library keeps the transition rounds itself (device-side churn has no host trace)
        last_crash = sim.get_array(A.ARR_LAST_CRASH).astype(np.int64)
        last_rejoin = sim.get_array(A.ARR_LAST_REJOIN).astype(np.int64)
    else:
        last_crash, last_rejoin = event_history(events, n_total, round_now)
    observer_up = alive[first:first + n_local][:, None]
    occupied = (st != A.VACANT) & observer_up
    member = np.where(occupied, nbr, 0).astype(np.int64)
    m_up = alive[member]
    down = occupied & ~m_up
    up = occupied & m_up
    dead = st == A.DEAD
    crash_at = last_crash[member]
    detected = down & dead & (vlast >= crash_at)
    lat = np.clip((vlast - crash_at)[detected], 0, max_latency)
    out = {
        "entries": int(occupied.sum()),
        "down_entries": int(down.sum()),
        "detected": int(detected.sum()),
        "stale_dead": int((down & dead & (vlast < crash_at)).sum()),
        "undetected": int((down & ~dead).sum()),
        "suspected": int((up & (st == A.SUSPECT)).sum()),
        "false_dead": int((up & dead & (vlast > last_rejoin[member])).sum()),
        "stale_dead_up": int((up & dead & (vlast <= last_rejoin[member])).sum()),
        "latency_hist": np.bincount(lat, minlength=max_latency + 1).astype(np.int64),
    }
    out["mismatches"] = out["undetected"] + out["suspected"] + out["false_dead"] + out["stale_dead_up"]
    return out


def merge_reports(reports):
    """Sum per-rank reports (all values are counts)."""
    out = {}
    for r in reports:
        for k, v in r.items():
            out[k] = out[k] + v if k in out else (v.copy() if isinstance(v, np.ndarray) else v)
    return out


def latency_stats(hist):
    """mean / p50 / p99 / max of a latency histogram (None when empty)."""
    total = int(hist.sum())
    if total == 0:
        return {"n": 0, "mean": None, "p50": None, "p99": None, "max": None}
    idx = np.arange(len(hist))
    cum = np.cumsum(hist)
    return {"n": total, "mean": float((idx * hist).sum() / total),
            "p50": int(np.searchsorted(cum, 0.5 * total)), "p99": int(np.searchsorted(cum, 0.99 * total)),
            "max": int(np.flatnonzero(hist)[-1])}


def run_sweep_point(sim, events, rounds, sample_every=10, reduce_sum=None):
    """Step `rounds` rounds, sampling the convergence count; returns the time series, the final counters and the final
    view report. `reduce_sum(list_of_numbers) -> list` adds values over ranks (identity for one GPU)."""
    red = reduce_sum or (lambda xs: xs)
    series = []
    done = 0
    device_ms = 0.0
    while done < rounds:
        step = min(sample_every, rounds - done)
        sim.step(step)
        if hasattr(sim, "last_step_ms"):
            device_ms += float(sim.last_step_ms())  # CUDA events around the call's kernels, this rank
        done += step
        series.append((done, int(red([sim.mismatches()])[0])))
    counters = [int(x) for x in red([int(v) for v in sim.counters()])]
    rep = view_report(sim, events, done)
    hist = np.array(red(rep.pop("latency_hist").tolist()), dtype=np.int64)
    keys = sorted(rep)
    vals = red([rep[k] for k in keys])
    rep = dict(zip(keys, (int(v) for v in vals)))
    rep["latency"] = latency_stats(hist)
    return {"mismatch_series": series, "counters": dict(zip(A.CTR_NAMES, counters)), "report": rep,
            "device_ms": device_ms, "device_us_per_round": device_ms * 1e3 / max(1, rounds)}
