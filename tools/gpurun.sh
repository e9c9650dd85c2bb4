#!/bin/bash
# Rebuild the in-tree .so (it travels with the snapshot), then run a command on a B200 box.
# usage: tools/gpurun.sh <timeout_s> '<command>' [extra gpurun flags]
set -e
cd "$(dirname "$0")/.."
python multinerf_b200/build.py > /dev/null
T=$1; shift
CMD=$1; shift
exec /usr/local/graft/bin/gpurun --timeout "$T" "$@" -- "$CMD"
