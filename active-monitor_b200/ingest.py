"""HealthCheck manifest -> classifier input (the step BEFORE the path).

In the deployed controller this conversion is the Go shim's (typed structs,
INTEGRATION.md section 2); this module does the same from a decoded YAML/JSON
document so that the reference's example manifests can be pushed through the
product's ladder (`am_healthcheck_classify`).  Field names are the JSON tags of
api/v1alpha1/healthcheck_types.go:32-66.
"""
from __future__ import annotations

import datetime as _dt

from .sweep import classify, remedy_is_empty


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
