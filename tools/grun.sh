#!/bin/bash
# usage: tools/grun.sh LOG TIMEOUT [--gpus N] -- 'command'   : gpurun with retries while the pod is busy (nothing is charged for those)
LOG=$1; TO=$2; shift 2
for i in $(seq 1 20); do
  /usr/local/graft/bin/gpurun --timeout $TO "$@" > $LOG 2>&1
  if grep -qE "status=transient|rc=3|no box|busy" $LOG && ! grep -q "status=ok" $LOG && ! grep -q "status=fail" $LOG; then sleep 120; else break; fi
done
tail -40 $LOG
