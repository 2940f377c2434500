#!/bin/bash
# tuning aid: CRC-32 tiles kernel duration vs tile size (ZIPC_B200_CRC_TILE; 0 = default, one tile per warp)
for t in ${@:-16384 32768 65536 131072 0}; do
  ZIPC_B200_CRC_TILE=$t timeout 200 python bench.py --no-also --steps 20 --warmup 3 2>&1 | tail -1 |
    python -c "import sys,json; d=json.loads(sys.stdin.read()); print('tile', $t, d['value'], d['roofline']['kernel_ms'], d['config']['crc32'])"
done
