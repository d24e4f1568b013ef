#!/bin/sh
# developer helper: retry a gpurun call while the pod answers "busy / draining" (exit code 3 or status=transient)
# usage: tools/gpurun_retry.sh <timeout-seconds> [--gpus N] -- '<command>'
T=$1; shift
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
	out=$(/usr/local/graft/bin/gpurun --timeout "$T" "$@" 2>&1)
	rc=$?
	echo "$out" | tail -80
	if echo "$out" | grep -q "status=transient\|no box or slot\|retry in a few minutes"; then
		sleep 120
		continue
	fi
	exit $rc
done
exit 3
