#!/bin/bash
# scripts/contact_ab.sh V1 V2 ...: contact-path timings of diffdart_b200/lib/variants/libddb_V.so ("base" = the product build):
# walker / half-cheetah forward step (scripts/contact_profile.py), box stacking (scripts/bench_contact.py), Atlas (scripts/atlas_bench.py)
for v in "$@"; do
  if [ $v = base ]; then L=$PWD/diffdart_b200/lib/libddb.so; else L=$PWD/diffdart_b200/lib/variants/libddb_$v.so; fi
  echo "== $v"
  DDB_LIB=$L timeout 300 python scripts/contact_profile.py 2>&1 | tail -2 | cut -c1-48
  DDB_LIB=$L timeout 300 python scripts/bench_contact.py 16384 3 2>&1 | cut -c50-135
  DDB_LIB=$L timeout 300 python scripts/atlas_bench.py 4096 5 2>&1 | tail -1 | cut -c80-150
done
