#!/bin/bash
# compute-sanitizer passes over the smoke workload (fused FC kernel + CNN front-end, 1013 images each) and the option paths
mkdir -p gpurun_out
for tool in memcheck synccheck; do
  echo "== compute-sanitizer --tool $tool"
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 5 python -c "
import __graft_entry__ as g
g.smoke()
import numpy as np
from bitnetmcu_b200 import engine as E, _lib
from bitnetmcu_b200.model import Model
x = np.random.default_rng(0).normal(size=(333, 256)).astype(np.float32)
q = E.quantize_images(x)
e = E.Engine(Model.load('tests/golden/models/binary160.bnm'))
e.set_option(_lib.OPT_PATH, _lib.PATH_LAYERS); a = e.infer(q)
e.set_option(_lib.OPT_PATH, _lib.PATH_TCGEN05); b = e.infer(q)
assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
print('layers == tcgen05 on binary160, quantised input ok')
" > gpurun_out/sanitize_$tool.log 2>&1
  echo "exit $?"; grep -E "ERROR SUMMARY|smoke|ok$|Error|error" gpurun_out/sanitize_$tool.log | head -12
done
