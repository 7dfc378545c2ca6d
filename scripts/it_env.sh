#!/usr/bin/env bash
# Integration-test environment: a stand-alone shard-server group on this machine.
#
# Counterpart of the reference's spark-test-env.sh (ENVSH:14-84), which creates / starts / stops one
# docker container with Spark + HDFS and `exec`s spark-submit inside it.  Here the "environment" is a
# long-lived separate server group (cf. `spark-submit --class glint.Main ... spark -c separate-glint.conf`,
# SBT:51-59) listening on 127.0.0.1:$PORT, tracked by a pid file.
#
#   scripts/it_env.sh start [num_servers]   start the group in the background, wait until it is ready
#   scripts/it_env.sh status                print the master address or "stopped"
#   scripts/it_env.sh exec <cmd...>         run a command with GW2V_IT_SERVER_HOST exported
#   scripts/it_env.sh stop                  terminate the group (exact pid only)
#   scripts/it_env.sh rm                    stop + delete state and logs
set -euo pipefail
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
STATE="${GW2V_IT_STATE:-$ROOT/.it_env}"
PORT="${GW2V_IT_PORT:-13380}"
PY="${PY:-python}"
mkdir -p "$STATE"

running() { [[ -f "$STATE/pid" ]] && kill -0 "$(cat "$STATE/pid")" 2>/dev/null; }

case "${1:-}" in
  start)
    if running; then echo "already running: $(cat "$STATE/master")"; exit 0; fi
    n="${2:-2}"
    rm -f "$STATE/ready"
    # per-deployment secret shared by the group and its clients (parallel/wire.py); 0600
    ( umask 077; head -c 32 /dev/urandom | od -An -tx1 | tr -d ' \n' > "$STATE/secret" )
    export GW2V_SERVER_SECRET_FILE="$STATE/secret"
    ( cd "$ROOT" && exec setsid "$PY" -m glint_word2vec_b200.parallel.server --num-servers "$n" --port "$PORT" \
        --bind 127.0.0.1 -c "${GW2V_IT_CONF:-$ROOT/configs/it-separate-server.json}" --device "${GW2V_IT_DEVICE:-cpu}" --ready-file "$STATE/ready" \
        > "$STATE/server.log" 2>&1 ) &
    echo $! > "$STATE/pid"
    for _ in $(seq 1 240); do
      [[ -f "$STATE/ready" ]] && break
      running || { echo "server group died, see $STATE/server.log" >&2; tail -20 "$STATE/server.log" >&2; exit 1; }
      sleep 0.5
    done
    [[ -f "$STATE/ready" ]] || { echo "server group did not become ready" >&2; exit 1; }
    echo "127.0.0.1:$PORT" > "$STATE/master"
    echo "master = 127.0.0.1:$PORT"
    ;;
  status)
    if running; then cat "$STATE/master"; else echo stopped; fi
    ;;
  exec)
    shift
    running || { echo "environment is not running (scripts/it_env.sh start)" >&2; exit 1; }
    GW2V_SERVER_SECRET_FILE="$STATE/secret" GW2V_IT_SERVER_HOST="$(cat "$STATE/master")" "$@"
    ;;
  stop)
    if running; then
      pid="$(cat "$STATE/pid")"
      kill -- "-$pid" 2>/dev/null || kill "$pid" 2>/dev/null || true    # the process group we started, nothing else
      for _ in $(seq 1 40); do running || break; sleep 0.25; done
    fi
    rm -f "$STATE/pid" "$STATE/master" "$STATE/ready"
    echo stopped
    ;;
  rm)
    "$0" stop >/dev/null
    rm -rf "$STATE"
    ;;
  *)
    sed -n 2,16p "$0"; exit 2
    ;;
esac
