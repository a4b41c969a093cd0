"""HealthCheck manifest -> classifier input (the step BEFORE the path).

In the deployed controller this conversion is the Go shim's (typed structs,
INTEGRATION.md section 2); this module does the same from a decoded YAML/JSON
document so that the reference's example manifests can be pushed through the
product's ladder (`am_healthcheck_classify`).  Field names are the JSON tags of
api/v1alpha1/healthcheck_types.go:32-66.
"""
from __future__ import annotations

import ctypes as C
import datetime as _dt

import numpy as np

from . import _lib as L
from .sweep import AmError, _as_bytes, classify, remedy_is_empty


def _unix(ts):
    """metav1.Time as serialised by the API server (RFC 3339, second precision)."""
    if ts is None:
        return None
    if isinstance(ts, (int, float)):
        return int(ts)
    if isinstance(ts, _dt.datetime):
        d = ts if ts.tzinfo else ts.replace(tzinfo=_dt.timezone.utc)
        return int(d.timestamp())
    return int(_dt.datetime.fromisoformat(str(ts).replace("Z", "+00:00")).timestamp())


def healthcheck_kwargs(doc: dict) -> dict:
    """keyword arguments of `classify()` for one HealthCheck document."""
    spec = doc.get("spec") or {}
    status = doc.get("status") or {}
    wf = spec.get("workflow") or {}
    rw = spec.get("remedyworkflow") or {}
    # RemedyWorkflow.IsEmpty (healthcheck_types.go:104-106): DeepEqual with the zero value
    has_remedy = not remedy_is_empty(rw.get("generateName") or "", rw.get("resource") is None,
                                     int(rw.get("workflowtimeout") or 0), rw.get("rbacRules") is None)
    return dict(
        repeat_after_sec=int(spec.get("repeatAfterSec") or 0),
        cron=(spec.get("schedule") or {}).get("cron") or "",
        has_resource=wf.get("resource") is not None,            # hcc.go:227
        has_remedy=has_remedy,
        remedy_runs_limit=int(spec.get("remedyRunsLimit") or 0),
        remedy_reset_interval=int(spec.get("remedyResetInterval") or 0),
        finished_at=_unix(status.get("finishedAt")),
        remedy_finished_at=_unix(status.get("remedyFinishedAt")),
        success_count=int(status.get("successCount") or 0),
        failed_count=int(status.get("failedCount") or 0),
        remedy_success_count=int(status.get("remedySuccessCount") or 0),
        remedy_failed_count=int(status.get("remedyFailedCount") or 0),
        remedy_total_runs=int(status.get("remedyTotalRuns") or 0),
    )


def record_from_manifest(doc: dict):
    """(rc, packed record) for one HealthCheck document."""
    return classify(**healthcheck_kwargs(doc))


_KW = ("repeat_after_sec", "cron", "has_resource", "has_remedy", "remedy_runs_limit",
       "remedy_reset_interval", "finished_at", "remedy_finished_at", "success_count", "failed_count",
       "remedy_success_count", "remedy_failed_count", "remedy_total_runs")


def classify_batch(items, n_threads: int = 0):
    """`am_healthcheck_classify_batch` over a list of `classify()` keyword dicts (or manifests).

    Returns (rc array int32[n], records RECORD_DTYPE[n]).  The ladder runs in the C++
    classifier on `n_threads` host threads (0 = all); this function only lays the inputs out
    as the am_healthcheck_t array a cgo caller would pass.
    """
    kws = [healthcheck_kwargs(it) if ("spec" in it or "status" in it) else it for it in items]
    n = len(kws)
    arr = (L.AmHealthCheck * n)()
    keep = []  # the cron bytes must outlive the call (the library copies nothing before it returns)
    for i, kw in enumerate(kws):
        unknown = set(kw) - set(_KW) - {"fail_p8", "timer_armed"}
        if unknown:
            raise TypeError(f"item {i}: unknown fields {sorted(unknown)}")
        cron = kw.get("cron", "") or ""
        raw = bytes(_as_bytes(cron))
        keep.append(raw)
        h = arr[i]
        h.repeat_after_sec = int(kw.get("repeat_after_sec", 0) or 0)
        h.cron, h.cron_len = raw, len(raw)
        h.has_resource = int(bool(kw.get("has_resource", True)))
        h.has_remedy = int(bool(kw.get("has_remedy", False)))
        h.remedy_runs_limit = int(kw.get("remedy_runs_limit", 0) or 0)
        h.remedy_reset_interval = int(kw.get("remedy_reset_interval", 0) or 0)
        fa, rfa = kw.get("finished_at"), kw.get("remedy_finished_at")
        h.finished_at, h.finished_at_set = (int(fa), 1) if fa is not None else (0, 0)
        h.remedy_finished_at, h.remedy_finished_at_set = (int(rfa), 1) if rfa is not None else (0, 0)
        h.success_count = int(kw.get("success_count", 0) or 0)
        h.failed_count = int(kw.get("failed_count", 0) or 0)
        h.remedy_success_count = int(kw.get("remedy_success_count", 0) or 0)
        h.remedy_failed_count = int(kw.get("remedy_failed_count", 0) or 0)
        h.remedy_total_runs = int(kw.get("remedy_total_runs", 0) or 0)
        h.fail_p8 = int(kw.get("fail_p8", 0) or 0)
        h.timer_armed = int(bool(kw.get("timer_armed", True)))  # same default as sweep.classify()
    recs = np.zeros(n, dtype=L.RECORD_DTYPE)
    rcs = np.zeros(n, dtype=np.int32)
    bad = L.u64(0)
    rc = L.load().am_healthcheck_classify_batch(C.cast(arr, C.c_void_p), n, recs.ctypes.data,
                                                rcs.ctypes.data, n_threads, C.byref(bad))
    if rc != L.AM_OK:
        raise AmError(rc, "am_healthcheck_classify_batch")
    assert int(bad.value) == int(np.count_nonzero(rcs))
    del keep
    return rcs, recs


def ingest_json(text, n_threads: int = 0, timer_armed: bool = True):
    """`am_healthcheck_ingest_json`: a JSON document (one HealthCheck, an array, or a List with
    "items") -> (rc array int32[n], records RECORD_DTYPE[n]).  Field extraction and the ladder both
    run natively, on `n_threads` host threads (0 = all)."""
    raw = text if isinstance(text, (bytes, bytearray)) else text.encode("utf-8")
    lib = L.load()
    n = L.u64(0)
    rc = lib.am_healthcheck_ingest_json(raw, len(raw), int(bool(timer_armed)), None, None, 0, C.byref(n), n_threads)
    if rc not in (L.AM_OK, L.AM_E_NOSPACE):
        raise AmError(rc, "am_healthcheck_ingest_json")
    recs = np.zeros(n.value, dtype=L.RECORD_DTYPE)
    rcs = np.zeros(n.value, dtype=np.int32)
    if n.value:
        rc = lib.am_healthcheck_ingest_json(raw, len(raw), int(bool(timer_armed)), recs.ctypes.data, rcs.ctypes.data,
                                            n.value, C.byref(n), n_threads)
        if rc != L.AM_OK:
            raise AmError(rc, "am_healthcheck_ingest_json")
    return rcs, recs
