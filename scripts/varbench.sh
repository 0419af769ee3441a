#!/bin/bash
# scripts/varbench.sh V1 V2 ...: kbench timings of diffdart_b200/lib/variants/libddb_V.so ("base" = the product build)
for v in "$@"; do
  if [ $v = base ]; then L=$PWD/diffdart_b200/lib/libddb.so; else L=$PWD/diffdart_b200/lib/variants/libddb_$v.so; fi
  echo -n "$v: "; DDB_LIB=$L python scripts/kbench.py --reps 400 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(' '.join('%s=%.1f'%(k,v) for k,v in d.items() if k.endswith('_us') or 'lanes' in k))"
done
