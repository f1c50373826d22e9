#!/usr/bin/env bash
# gpurun with retries while the pod answers "busy / backing off" (exit code 3, nothing charged).
# usage: tools/gpurun_retry.sh [--gpus N] [--timeout S] -- '<command>'
for attempt in 1 2 3 4 5 6 7 8 9 10 11 12; do
  /usr/local/graft/bin/gpurun "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  echo "[gpurun_retry] attempt $attempt answered busy; sleeping 120 s"
  sleep 120
done
exit 3
