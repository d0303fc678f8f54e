#!/bin/bash
# compute-sanitizer passes over the smoke workload (fused FC kernel + CNN front-end, 1013 images each) and the option paths
mkdir -p gpurun_out
for tool in memcheck synccheck racecheck; do
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
# four-warpgroup form (2bitsym-96) and the shared-memory-activation form (binary160 above) over several tiles per CTA
import numpy as _np
for nm in ('2bitsym96', 'binary160'):
    e4 = E.Engine(Model.load('tests/golden/models/%s.bnm' % nm))
    big = _np.random.default_rng(5).integers(-128, 128, size=(148 * 128 * 2 + 9, 256)).astype(_np.int8)
    e4.set_option(_lib.OPT_PATH, _lib.PATH_LAYERS); a4 = e4.infer(big)
    e4.set_option(_lib.OPT_PATH, _lib.PATH_TCGEN05); b4 = e4.infer(big)
    assert _np.array_equal(a4[0], b4[0]) and _np.array_equal(a4[1], b4[1])
    e4.close()
print('four-warpgroup and smem-activation forms ok')
# both CNN front-end kernels on a 48-channel model (tiles that straddle images), emulation mode, fused gather on one GPU
import torch, ctypes as C
m48 = Model.load('tests/golden/models/cnn_48.bnm')
imgs = np.random.default_rng(1).integers(-128, 128, size=(777, 256)).astype(np.int8)
outs = []
for fe in (_lib.CNN_CUDA_CORES, _lib.CNN_TENSOR_CORES):
    e = E.Engine(m48); e.set_option(_lib.OPT_CNN_FRONTEND, fe); outs.append(e.infer(imgs)); e.close()
assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
e = E.Engine(Model.load('tests/golden/models/fc.bnm'))
lg = e.inference_quantized(x)
n = 148 * 128 + 77
d_img = torch.randint(-128, 128, (n, 256), dtype=torch.int8, device='cuda')
d_log = torch.empty((n, 10), dtype=torch.int32, device='cuda'); d_lab = torch.empty(n, dtype=torch.int32, device='cuda')
g_log = torch.zeros((n + 256, 10), dtype=torch.int32, device='cuda'); g_lab = torch.zeros(n + 256, dtype=torch.int32, device='cuda')
g = _lib.BnmGather(); g.n_labels_dst = 1; g.n_logits_dst = 1; g.row_offset = 128
g.labels_dst[0] = g_lab.data_ptr(); g.logits_dst[0] = g_log.data_ptr()
_lib.check(e.lib.bnm_infer_batch_device_gather(e.handle, C.c_void_p(d_img.data_ptr()), n, C.c_void_p(d_log.data_ptr()), C.c_void_p(d_lab.data_ptr()), C.byref(g), None), 'gather')
torch.cuda.synchronize()
assert torch.equal(g_log[128:128 + n], d_log) and torch.equal(g_lab[128:128 + n], d_lab)
print('cnn front-ends agree, emulation mode ran, fused gather ok')
# float32 images in, scaling fused into the FC kernel's load stage (quantiser warps write the swizzled int8 stage by hand)
nf = 148 * 128 * 2 + 53
xf = torch.randn((nf, 256), dtype=torch.float32, device='cuda') * 3
f_log = torch.empty((nf, 10), dtype=torch.int32, device='cuda'); f_lab = torch.empty(nf, dtype=torch.int32, device='cuda')
e.infer_device_f32(xf, f_log, f_lab)
torch.cuda.synchronize()
ref = e.infer(E.quantize_images(xf.cpu().numpy()))
assert np.array_equal(f_log.cpu().numpy(), ref[0])
print('fused float input ok')
" > gpurun_out/sanitize_$tool.log 2>&1
  echo "exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke|ok$|Error|error" gpurun_out/sanitize_$tool.log | head -12
done
