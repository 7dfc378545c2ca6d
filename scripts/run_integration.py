#!/usr/bin/env python
"""Integration-test runner: the 15 acceptance scenarios, exit status 1 on any failure.

Counterpart of the reference's IT ``Main`` (ITMAIN:8-15), which runs the ScalaTest ``Runner`` inside
``spark-submit`` and exits 1 on failure, and of the custom ``it:test`` sbt task (SBT:42-82) that first
brings up the environment, launches the separate parameter-server application, runs the suite and
tears everything down.

    python scripts/run_integration.py            # starts scripts/it_env.sh, runs the suite, stops it
    python scripts/run_integration.py --no-env   # environment managed by the caller
"""
import argparse
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--no-env", action="store_true", help="do not start/stop the separate server group")
    ap.add_argument("--servers", type=int, default=2)
    ap.add_argument("--gpu", action="store_true", help="also run the GPU tier")
    ap.add_argument("pytest_args", nargs="*")
    args = ap.parse_args()
    env_sh = os.path.join(ROOT, "scripts", "it_env.sh")
    env = dict(os.environ)
    started = False
    try:
        if not args.no_env:
            subprocess.check_call([env_sh, "start", str(args.servers)])
            started = True
            env["GW2V_IT_SERVER_HOST"] = subprocess.check_output([env_sh, "status"], text=True).strip()
        cmd = [sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_golden_spec.py"), "-q", "-x"]
        if not args.gpu:
            cmd += ["-m", "not gpu"]
        rc = subprocess.call(cmd + args.pytest_args, cwd=ROOT, env=env)
    finally:
        if started:
            subprocess.call([env_sh, "stop"])
    return 1 if rc != 0 else 0


if __name__ == "__main__":
    sys.exit(main())
