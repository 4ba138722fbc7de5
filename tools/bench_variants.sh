#!/bin/bash
# usage: [VAR=OTB_LIB|OTB_SPEC_LIB] [ARGS="--engine specialised"] tools/bench_variants.sh lib1.so lib2.so ...
# prints ms/step, trace kernel ms, roofline frac per engine build (OTB_LIB: whole engine; OTB_SPEC_LIB: specialised variant)
VAR=${VAR:-OTB_LIB}
for lib in "$@"; do
  env $VAR=$lib python bench.py --steps 5 --warmup 3 --no-cpu --no-compare $ARGS 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$lib', 'step_ms=%.3f trace_ms=%.3f frac=%.4f e2e_ms=%.3f' % (d['ms_per_step'], d['config']['trace_kernel_ms'], d['roofline']['frac'], d['e2e']['ms_per_step']))"
done
