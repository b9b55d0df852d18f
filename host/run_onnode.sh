#!/bin/bash
# Starts one dorylus_b200_run process per GPU of this box -- the stand-in for run/run-onnode, which
# starts one graph server per machine.  The ranks meet through a fresh rendezvous directory.
#   host/run_onnode.sh N --datasetdir D/ --featuresfile F --labelsfile L --layerfile C [driver flags]
set -e
N=$1; shift
HERE=$(cd "$(dirname "$0")" && pwd)
BIN=${DORY_RUN_BIN:-$HERE/dorylus_b200_run}
RDV=$(mktemp -d /tmp/dory_rendezvous.XXXXXX)
pids=()
for ((i = 0; i < N; ++i)); do
    "$BIN" "$@" --numnodes "$N" --nodeid "$i" --device "$i" --rendezvous "$RDV" &
    pids+=($!)
done
rc=0
for p in "${pids[@]}"; do wait "$p" || rc=$?; done
rm -rf "$RDV"
exit $rc
