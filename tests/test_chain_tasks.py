"""Round-1 task derivation (Appendix C) against a per-chain loop written from the reference's control flow."""
import numpy as np

from lordfast_b200 import api, sim
from lordfast_b200.chain_tasks import KIND_GAP, KIND_HEAD, KIND_TAIL, workload_tasks


def test_round1_tasks_match_per_chain_loop():
    w = sim.make_workload(150_000, 30, 3000, 0.12, 0.15, seed=4, sv_frac=0.4)
    tasks, chain, kind = workload_tasks(w)
    exp = []
    for i in range(w.n_reads):
        s = w.chain(i).astype(np.int64)
        L = int(w.read_off[i + 1] - w.read_off[i])
        fl = api.LF_F_READ_REV if w.is_rev[i] else 0
        a = int(s[0, 1])
        if a > 0 and s[0, 0] - (a + 20) >= 0:
            exp.append((i, 0, a, int(s[0, 0]) - a - 20, a + 20, fl | api.LF_F_REVERSE_BOTH | (api.LF_F_NO_PATH if a > 500 else 0), api.LF_MODE_SHW, KIND_HEAD))
        for k in range(len(s) - 1):
            qs, ts = int(s[k, 1] + s[k, 2]), int(s[k, 0] + s[k, 2])
            ql, tl = int(s[k + 1, 1]) - qs, int(s[k + 1, 0]) - ts
            if ql > 0 and tl > 0:
                exp.append((i, qs, ql, ts, tl, fl, api.LF_MODE_NW, KIND_GAP))
        qs = int(s[-1, 1] + s[-1, 2]); b = L - qs; ts = int(s[-1, 0] + s[-1, 2])
        if b > 0 and ts + b + 20 - 1 <= len(w.ref) - 1:
            exp.append((i, qs, b, ts, b + 20, fl | (api.LF_F_NO_PATH if b > 500 else 0), api.LF_MODE_SHW, KIND_TAIL))
    assert len(exp) == len(tasks)
    assert any(e[5] & api.LF_F_NO_PATH for e in exp)   # the SV mix holds junk heads / tails above _pf_clipLen
    for e, t, c, k in zip(exp, tasks, chain, kind):
        assert e == (int(t["read_id"]), int(t["q_off"]), int(t["q_len"]), int(t["t_off"]), int(t["t_len"]), int(t["flags"]), int(t["mode"]), int(k))
        assert int(c) == e[0]
