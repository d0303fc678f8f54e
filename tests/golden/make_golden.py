#!/usr/bin/env python
"""Generate tests/golden/ from the reference itself.  Run ONLY in the build container (needs /root/reference):

    python tests/golden/make_golden.py

What it writes (all small, committed, and read at test time on the GPU box where /root/reference is absent):

  models/<name>.bnm   every shipped reference model header converted to the BNM1 blob by our header parser,
                      plus models for encodings that ship only as float checkpoints (Binary-160, 4bit, 2bitsym-96:
                      quantised with the reference's own BitLinear.weight_quant imported from /root/reference and
                      packed by bitnetmcu_b200.pack) and Ternary-64, plus random-code
                      models for 8bit / NF4 / Binary / FP130 (every bit pattern is a legal weight).
  digits.npz          the 10 MNIST test digits + labels of BitNetMCU_MNIST_test_data.h
  golden.npz          per model: int32 logits + labels on the 10 digits and on 256 xorshift32(seed 12345) images,
                      computed by the UNMODIFIED reference kernels (oracle/_ref/libbitnetmcu_ref.so)
  kat.json            label KAT of the reference DLLs (Inference()), CRC / label-sum KATs of SURVEY.md 8c

The oracle (oracle/bitnet_oracle.c) is pinned against these by tests/test_oracle_golden.py.
"""
import json
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))

from bitnetmcu_b200 import model as M  # noqa: E402
from bitnetmcu_b200 import pack as P  # noqa: E402
from oracle import oracle as O  # noqa: E402


def checkpoint_fc_model(ckpt, quant_type, widths, name):
    """Quantise a float checkpoint with the reference's own BitLinear.weight_quant (BitNetMCU.py:131-180)."""
    import torch
    sys.path.insert(0, REF)
    from BitNetMCU import BitLinear  # the reference's Python, imported in place
    sd = torch.load(os.path.join(REF, "modeldata", ckpt), map_location="cpu")
    keys = ["fc1.weight", "fc2.weight", "fc3.weight", "fcl.weight"]
    if "model.1.weight" in sd:
        keys = ["model.1.weight", "model.3.weight", "model.fc3.weight", "classifier.weight"]
    layers = []
    for i, k in enumerate(keys):
        w = sd[k].float()
        bl = BitLinear(w.shape[1], w.shape[0], QuantType=quant_type)
        skey = k.replace(".weight", ".s")
        if skey in sd:
            bl.s = torch.nn.Parameter(sd[skey].float(), requires_grad=False)
        else:  # old checkpoints carry no clipping scalar: proportional default (BitNetMCU.py:104-108)
            bl.update_clipping_scalar(w, "prop", 0.25)
        u, _, _ = bl.weight_quant(w)
        layers.append(P.fc_layer_from_levels(f"L{i + 1}", quant_type, u.numpy()))
        assert tuple(w.shape) == (widths[i + 1], widths[i])
    m = M.Model(model_class=M.MODEL_FCMNIST, layers=layers, source=f"{ckpt}:{quant_type}")
    m.validate()
    return name, m


def main():
    O.build(quiet=True)
    ref, orc = O.Reference(), O.Oracle()
    os.makedirs(os.path.join(OUT, "models"), exist_ok=True)

    models = {}
    shipped = {"fc": "BitNetMCU_model_fc.h", "cnn": "BitNetMCU_model_cnn.h", "12k": "mcu/BitNetMCU_model_12k.h",
               "12k_FP130": "mcu/BitNetMCU_model_12k_FP130.h", "1k": "mcu/BitNetMCU_model_1k.h",
               "cnn_16": "mcu/BitNetMCU_model_cnn_16.h", "cnn_16small": "mcu/BitNetMCU_model_cnn_16small.h",
               "cnn_32": "mcu/BitNetMCU_model_cnn_32.h", "cnn_48": "mcu/BitNetMCU_model_cnn_48.h",
               "cnn_letters": "mcu/BitNetMCU_model_cnn_letters.h"}
    for name, h in shipped.items():
        models[name] = M.parse_header(os.path.join(REF, h))

    for name, m in [
        checkpoint_fc_model("a11_Opt12k_cos_Aug_BitMnist_PerTensor_Binary_RMS_width160_160_160_lr0.001_decay0.1_stepsize10_bs128_epochs60.pth",
                            "Binary", (256, 160, 160, 160, 10), "binary160"),
        checkpoint_fc_model("opt_Cosine_lr0.001_Aug_BitMnist_PerTensor_4bit_RMS_width64_64_64_bs128_epochs60.pth",
                            "4bit", (256, 64, 64, 64, 10), "4bit64"),
        checkpoint_fc_model("a11_Opt12k_cos_Aug_BitMnist_PerTensor_2bitsym_RMS_width96_96_96_lr0.001_decay0.1_stepsize10_bs128_epochs60.pth",
                            "2bitsym", (256, 96, 96, 96, 10), "2bitsym96"),
        checkpoint_fc_model("opt_Cosine_lr0.001_Aug_BitMnist_PerTensor_4bitsym_RMS_width64_64_64_bs128_epochs60.pth",
                            "Ternary", (256, 64, 64, 64, 10), "ternary64"),
        checkpoint_fc_model("opt_Cosine_lr0.001_Aug_BitMnist_PerTensor_4bitsym_RMS_width64_64_64_bs128_epochs60.pth",
                            "8bit", (256, 64, 64, 64, 10), "8bit64"),
    ]:
        models[name] = m
    models["rand_binary64"] = P.random_fc_model(M.ENC_BINARY, seed=1)
    models["rand_8bit64"] = P.random_fc_model(M.ENC_8BIT, seed=2)
    models["rand_fp130_64"] = P.random_fc_model(M.ENC_FP130, seed=3)   # contains +128 and -128 weights
    models["rand_nf4_64"] = P.random_fc_model(M.ENC_NF4, seed=4)       # reference decodes id 36 to zeros
    models["rand_ternary64"] = P.random_fc_model(M.ENC_TERNARY, seed=5)

    digits, labels = M.parse_test_data_header(os.path.join(REF, "BitNetMCU_MNIST_test_data.h"))
    np.savez_compressed(os.path.join(OUT, "digits.npz"), images=digits, labels=labels)
    xs = orc.xorshift_images(256, 256, 12345)

    gold = {}
    for name, m in models.items():
        m.save(os.path.join(OUT, "models", name + ".bnm"))
        ld, lab_d = ref.infer(m, digits, threads=1)
        lx, lab_x = ref.infer(m, xs, threads=1)
        gold[name + "/digits_logits"], gold[name + "/digits_labels"] = ld, lab_d
        gold[name + "/xs_logits"], gold[name + "/xs_labels"] = lx, lab_x
        print(f"{name:16s} {m.describe()[:70]:70s} digit labels {lab_d.tolist()}")
    np.savez_compressed(os.path.join(OUT, "golden.npz"), **gold)

    # KATs of SURVEY.md 8c, re-derived here from the reference
    fc, cnn = models["fc"], models["cnn"]
    x1k = orc.xorshift_images(1000)
    l1k, lab1k = ref.infer(fc, x1k)
    x400k = orc.xorshift_images(400000)
    _, lab400k = ref.infer(fc, x400k)
    _, lab40k = ref.infer(cnn, x400k[:40000])
    kat = {
        "reference_labels": labels.tolist(),
        "dll_fc_labels": O.reference_dll_labels("fc", digits).tolist(),
        "dll_cnn_labels": O.reference_dll_labels("cnn", digits).tolist(),
        "xorshift_first_pixels": x1k[0, :8].tolist(),
        "fc_xs_logits0": l1k[0].tolist(),
        "fc_xs_first8_labels": lab1k[:8].tolist(),
        "fc_xs_crc32_first1000": "%08x" % zlib.crc32(l1k.astype("<i4").tobytes()),
        "fc_xs_sum_labels_1000": int(lab1k.sum()),
        "fc_xs_sum_labels_400000": int(lab400k.sum()),
        "cnn_xs_sum_labels_40000": int(lab40k.sum()),
    }
    with open(os.path.join(OUT, "kat.json"), "w") as f:
        json.dump(kat, f, indent=1)
    print(json.dumps(kat))


if __name__ == "__main__":
    main()
