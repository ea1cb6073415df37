#!/bin/bash
# tools/gpurun_retry.sh <timeout> <command...>: gpurun answers "transient" (exit 3, nothing charged) while the pod drains its GPU slots;
# retry every two minutes, for at most an hour
T=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout $T "$@" > /tmp/gpurun_retry.$$ 2>&1
  rc=$?
  if ! grep -q "status=transient" /tmp/gpurun_retry.$$; then cat /tmp/gpurun_retry.$$; rm -f /tmp/gpurun_retry.$$; exit $rc; fi
  sleep 120
done
cat /tmp/gpurun_retry.$$; rm -f /tmp/gpurun_retry.$$
exit 3
