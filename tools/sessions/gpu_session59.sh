#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/s59
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 900 python -m pytest tests/test_frames.py -x -q -m gpu 2>&1 | tail -5
