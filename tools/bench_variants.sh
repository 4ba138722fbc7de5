#!/bin/bash
# usage: tools/bench_variants.sh lib1.so lib2.so ...   (prints ms/step, trace kernel ms, roofline frac per engine build)
for lib in "$@"; do
  OTB_LIB=$lib python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$lib', 'step_ms=%.3f trace_ms=%.3f frac=%.4f e2e_ms=%.3f' % (d['ms_per_step'], d['config']['trace_kernel_ms'], d['roofline']['frac'], d['e2e']['ms_per_step']))"
done
