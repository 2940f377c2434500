#!/bin/bash
# tuning aid: bench the deflate workload with alternative builds of the library (zipc_b200/libzipc_b200_<tag>.so)
for tag in "$@"; do
  lib=""; [ "$tag" != default ] && lib=$PWD/zipc_b200/libzipc_b200_$tag.so
  ZIPC_B200_LIB=$lib timeout 400 python bench.py --workload deflate --steps 3 --warmup 3 2>&1 | tail -1 |
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$tag', d['value'], d['roofline']['kernel_ms'], d['e2e']['value'], d['config'].get('ratio'))"
done
