"""Comparison of a device result (pandaseq_b200.Context.assemble_host / assemble_device output) with a
checker result (oracle_lib.assemble).  Integer fields and merged bases must be identical; floating
point fields within TOL (BASELINE.json north_star: 1e-6)."""
from __future__ import annotations

import numpy as np

TOL = 1e-6
INT_FIELDS = ("overlap", "seq_len", "mismatches", "degenerates", "examined", "fwd_offset", "rev_offset")


def compare(got: dict, want: dict, *, emitted_only_ok=False, check_counters=True) -> dict:
    """got: dict(results=structured array, seq_nt, seq_p, counters); want: oracle dict.
    emitted_only_ok: the checker only fills fields for status OK rows (the compiled reference);
    otherwise OK and LOWQ rows are compared (the oracle port fills both)."""
    res = got["results"]
    n = len(res)
    rep = {"n": n, "ok": True, "max_dq": 0.0, "max_dp": 0.0, "max_dbase_p": 0.0, "bad": {}}

    def fail(key, count):
        if count:
            rep["bad"][key] = int(count)
            rep["ok"] = False

    fail("status", (res["status"] != want["status"]).sum())
    fail("slow", (res["slow"] != want["slow"]).sum())
    if emitted_only_ok:
        rows = want["status"] == 0
    else:
        rows = (want["status"] == 0) | (want["status"] == 5) | (want["status"] >= 8)       # OK, LOWQ, rejected by a filter
    rows &= res["status"] == want["status"]
    for k in INT_FIELDS:
        if k == "examined" and not emitted_only_ok:
            fail(k, (res[k].astype(np.int64) != want[k]).sum())   # defined for every pair that reached align()
        else:
            fail(k, (res[k][rows].astype(np.int64) != want[k][rows]).sum())
    if rows.any():
        with np.errstate(invalid="ignore"):
            dq = np.abs(res["quality"][rows] - want["quality"][rows])
            dp = np.abs(res["est_prob"][rows] - want["est_prob"][rows])
            same_inf = np.isinf(res["est_prob"][rows]) & (res["est_prob"][rows] == want["est_prob"][rows])
            dp = np.where(same_inf, 0.0, dp)
        rep["max_dq"] = float(np.nanmax(dq))
        rep["max_dp"] = float(np.nanmax(dp))
        fail("quality", (~(dq <= TOL)).sum())
        fail("est_prob", (~(dp <= TOL)).sum())
        idx = np.nonzero(rows)[0]
        sl = want["seq_len"][idx]
        if got.get("seq_nt") is not None and want.get("seq_nt") is not None:
            w = min(got["seq_nt"].shape[1], want["seq_nt"].shape[1])
            mask = np.arange(w)[None, :] < sl[:, None]
            fail("seq_nt", ((got["seq_nt"][idx, :w] != want["seq_nt"][idx, :w]) & mask).sum())
            if got.get("seq_p") is not None and want.get("seq_p") is not None:
                d = np.abs(got["seq_p"][idx, :w] - want["seq_p"][idx, :w]) * mask
                rep["max_dbase_p"] = float(d.max()) if d.size else 0.0
                fail("seq_p", (d > TOL).sum())
    if check_counters and got.get("counters") is not None:
        fail("counters", (got["counters"] != want["counters"]).sum())
    return rep
