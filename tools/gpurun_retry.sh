#!/bin/bash
# usage: tools/gpurun_retry.sh [gpurun options] -- 'command'     retries while the pod has no free GPU slot (exit code 3)
for i in $(seq 1 40); do
    /usr/local/graft/bin/gpurun "$@"
    rc=$?
    if [ $rc -ne 3 ]; then exit $rc; fi
    echo "[gpurun_retry] no slot (attempt $i), retrying in 45 s" >&2
    sleep 45
done
exit 3
