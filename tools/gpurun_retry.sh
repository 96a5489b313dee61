#!/bin/bash
# gpurun_retry.sh [gpurun options] -- command : retries while the pod answers "busy" (exit code 3), at most 12 times
for i in $(seq 1 12); do
	/usr/local/graft/bin/gpurun "$@"
	rc=$?
	if [ $rc -ne 3 ]; then exit $rc; fi
	sleep 60
done
exit 3
