#!/usr/bin/env bash
# tools/gpurun_retry.sh <out-file> <gpurun args...> — retry a gpurun call while the pod answers "busy" (exit 3: nothing charged)
out=$1; shift
for i in $(seq 1 40); do
    /usr/local/graft/bin/gpurun "$@" > "$out" 2>&1
    rc=$?
    [ $rc -ne 3 ] && exit $rc
    sleep 90
done
exit 3
