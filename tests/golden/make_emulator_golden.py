"""Golden vectors of the reference's Python emulator for the GPU emulation mode.

Runs in THIS container only (imports /root/reference/BitNetMCU.py, which cannot travel to the GPU box):

    python tests/golden/make_emulator_golden.py   ->  tests/golden/emulator.npz

For each fixture model the unmodified ``QuantizedModel.inference_quantized`` (BitNetMCU.py:420-535) is fed a
``quantized_model`` list whose weight LEVELS are decoded from the packed fixture (integer weight / level scale: 4bitsym and
2bitsym levels are half-integers, BitNetMCU.py:155-164) and float32 images: the ten reference digits rescaled to floats and
random normal images.  Only encodings whose levels the packed integers determine exactly are used (not "4bit": its levels carry
a +0.01 offset, and not NF4).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")

MODELS = ["fc", "1k", "12k_FP130", "2bitsym96", "8bit64", "binary160", "ternary64", "cnn", "cnn_48", "cnn_16small"]
LEVEL_SCALE = {2: 2.0, 4: 2.0}


def quantized_model_list(m):
    from bitnetmcu_b200 import model as M, pack as P
    out = []
    width = m.img_bytes if m.model_class == 0 else m.channels * 4     # what the previous layer hands over
    for l in m.layers:
        if l.kind == M.LAYER_FC:
            # Ternary layers declare n_in padded to a multiple of 10 with zero trits (exportquant.py:132-137); the emulator sees the
            # unpadded float model, so the (all-zero) pad columns are dropped
            w = P.decode_int_weights(l).astype(np.float64)[:, :width] / LEVEL_SCALE.get(l.bitperweight, 1.0)
            width = l.n_out
            out.append({"layer_type": "BitLinear", "incoming_weights": l.n_in, "outgoing_weights": l.n_out,
                        "quantized_weights": w.tolist(), "WScale": "PerTensor", "bpw": 0, "quantization_type": M.ENC_NAMES[l.bitperweight]})
        elif l.kind == M.LAYER_CONV33:
            first = l.n_in == 16
            out.append({"layer_type": "BitConv2d", "in_channels": 1 if first else l.n_out, "out_channels": l.n_out, "kernel_size": (3, 3),
                        "stride": 1, "padding": 0, "groups": 1 if first else l.n_out, "quantized_weights": l.weights.astype(np.float64).tolist(),
                        "incoming_x": 0, "incoming_y": 0, "outgoing_x": 0, "outgoing_y": 0, "bpw": 8, "quantization_type": "8bit"})
        else:
            out.append({"layer_type": "MaxPool2d", "kernel_size": 2, "stride": 2})
    return out


def images():
    d = np.load(os.path.join(ROOT, "tests", "golden", "digits.npz"))
    rng = np.random.default_rng(2024)
    dig = d["images"].astype(np.float32) * np.float32(0.0247) + np.float32(0.01)     # de-quantised digits, arbitrary float scale
    rnd = rng.normal(size=(54, 256)).astype(np.float32)
    return np.concatenate([dig, rnd]).astype(np.float32)


def main():
    from BitNetMCU import QuantizedModel
    from bitnetmcu_b200.model import Model
    x = images()
    out = {"images": x}
    for name in MODELS:
        m = Model.load(os.path.join(ROOT, "tests", "golden", "models", name + ".bnm"))
        qm = QuantizedModel()
        qm.quantized_model = quantized_model_list(m)
        # the emulator scales per LAST axis (BitNetMCU.py:435): flattened images (batch, 256), the calling convention of
        # test_inference.py:153 and exportquant.py:547 (BitConv2d layers reshape them back, BitNetMCU.py:461-464)
        data = x
        with np.errstate(all="ignore"):
            logits = qm.inference_quantized(data)
        out[name] = np.asarray(logits, dtype=np.float64)
        print(name, out[name].shape, out[name][0][:4])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "emulator.npz"), **out)


if __name__ == "__main__":
    main()
