#!/bin/bash
# tuning aid: bench the inflate workload with alternative builds of the library (zipc_b200/libzipc_b200_<tag>.so)
for tag in "$@"; do
  lib=""; [ "$tag" != default ] && lib=$PWD/zipc_b200/libzipc_b200_$tag.so
  ZIPC_B200_LIB=$lib timeout 300 python bench.py --workload inflate --steps 5 --warmup 3 2>&1 | tail -1 |
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$tag', d['value'], d['roofline']['kernel_ms'], d['e2e']['value'])"
done
