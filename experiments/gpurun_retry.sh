#!/bin/bash
# usage: gpurun_retry.sh <log> <timeout> [--gpus N] -- <command>: retries while the pod answers "transient" (nothing charged)
LOG=$1; shift; TMO=$1; shift
EXTRA=""
if [ "$1" == "--gpus" ]; then EXTRA="--gpus $2"; shift; shift; fi
shift   # the "--"
for attempt in 1 2 3 4 5 6 7 8 9 10 11 12; do
  /usr/local/graft/bin/gpurun $EXTRA --timeout $TMO -- "$@" > $LOG 2>&1
  if grep -q "status=transient\|exit code 3\|no box or slot" $LOG; then sleep 150; continue; fi
  break
done
tail -150 $LOG | cut -c1-900
