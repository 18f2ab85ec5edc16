#!/bin/bash
# A/B of decoder-kernel variants (libraries under variants/, see tools/build_variant.sh): kernel ms per 4096 codewords
#   gpurun -- 'bash tools/ab_phi.sh base reg4 reg8 reg10'
for v in "$@"; do
  lib=""; [ "$v" != base ] && lib=$PWD/variants/$v.so
  SIONNA_B200_LIB=$lib python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-links --no-traffic 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$v', 'phi2dB %.3f ms'%r['kernel_ms'], ' '.join('%s %.3f'%(k[:18],v['kernel_ms']) for k,v in r['variants'].items()), 'ber', d['config']['ber']['bit_errors'])"
done
