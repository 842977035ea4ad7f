#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out/s65
mkdir -p $O
python -c "import __graft_entry__ as g; g.build()" > $O/build.log 2>&1
timeout 2400 python -m pytest tests -q -m gpu 2>&1 | tail -5
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
