#!/bin/bash
# ab.sh [variant ...]: stage table of bench.py (kernels only) for the in-tree build and each variants/<name> build
for v in main "$@"; do
  if [ "$v" = main ]; then unset MS_LIB_PATH; else export MS_LIB_PATH=/root/repo/variants/$v/libmicrosound_b200.so; fi
  python bench.py --steps 4 --warmup 3 --e2e-steps 0 --cpu-sample 0 ${RENDERS:+--renders $RENDERS} 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['ms_per_step'],2), {k:v['ms'] for k,v in d['stages'].items()})"
done
