"""Planner bindings for the tests: the product's (motion_primitive_library_b200.planner) plus the
TEST-ONLY harness oracle/liboracle_planner.so (same host A*, CPU-oracle env)."""
import subprocess
from pathlib import Path

from motion_primitive_library_b200.planner import (PlanArgs, PlanResult, QueryResult, Waypoint, load_fn,  # noqa: F401
                                                   make_args, plan, plan_batch, run_plan)

ROOT = Path(__file__).resolve().parent.parent


def plan_gpu(args):
    return plan(args)


def plan_oracle(args):
    p = ROOT / "oracle" / "liboracle_planner.so"
    if not p.exists():
        subprocess.check_call(["make", "-C", str(ROOT / "oracle"), "liboracle_planner.so"], stdout=subprocess.DEVNULL)
    lib, fn = load_fn(p, "orcp_plan")
    return run_plan(fn, lib, args)
