#!/bin/bash
set -u
cd "$(dirname "$0")/../.."
echo "== d=32"; timeout 120 python tools/prof_attn.py 32 8 56 3 tc 1 2>&1 | tail -9
echo "== d=8"; timeout 120 python tools/prof_attn.py 8 8 56 3 tc 1 2>&1 | tail -9
