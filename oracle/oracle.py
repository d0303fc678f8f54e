"""ctypes bindings for the CPU checker (TEST INFRASTRUCTURE -- never imported by bitnetmcu_b200/).

``Oracle``     -> oracle/libbnm_oracle.so      our C restatement (bitnet_oracle.c)
``Reference``  -> oracle/_ref/libbitnetmcu_ref.so  the unmodified reference kernels + ref_driver.c
Both expose ``infer(model, images, threads)`` -> (logits int32 [n, n_classes], labels uint32 [n])
and the four reference kernels one call at a time.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Tuple

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = "/root/reference"


class _Layer(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("bitperweight", C.c_int32), ("n_in", C.c_uint32), ("n_out", C.c_uint32),
                ("weights", C.c_void_p)]


def build(quiet: bool = True) -> None:
    """Compile the checker (and oracle/_ref when /root/reference is present).  Building the checker is not using it."""
    subprocess.run(["make", "-C", HERE], check=True, capture_output=quiet)


def _layer_table(model):
    keep = []
    arr = (_Layer * len(model.layers))()
    for i, l in enumerate(model.layers):
        ptr = None
        if l.weights is not None:
            w = np.ascontiguousarray(l.weights)
            keep.append(w)
            ptr = w.ctypes.data
        arr[i] = _Layer(l.kind, l.bitperweight, l.n_in, l.n_out, ptr)
    return arr, keep


def _as_images(images: np.ndarray, img_bytes: int) -> np.ndarray:
    a = np.ascontiguousarray(images, dtype=np.int8).reshape(-1, img_bytes)
    return a


class _Base:
    prefix = ""

    def __init__(self, path: str):
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing -- run `make -C oracle`")
        self.lib = C.CDLL(path)
        self.path = path

    # the four kernels, one call each (BitNetMCU_inference.h:15-60) ----------------------------
    def _fn(self, name):
        return getattr(self.lib, self.prefix + name)

    def relunorm(self, x: np.ndarray) -> Tuple[np.ndarray, int]:
        x = np.ascontiguousarray(x, dtype=np.int32)
        out = np.zeros(max(x.size, 1), dtype=np.int8)
        f = self._fn("ReLUNorm")
        f.restype = C.c_uint32
        pos = f(C.c_void_p(x.ctypes.data), C.c_void_p(out.ctypes.data), C.c_uint32(x.size))
        return out[: x.size], int(pos)

    def fclayer(self, act: np.ndarray, weights: np.ndarray, enc: int, n_in: int, n_out: int) -> np.ndarray:
        act = np.ascontiguousarray(act, dtype=np.int8)
        w = np.ascontiguousarray(weights)
        out = np.zeros(n_out, dtype=np.int32)
        f = self._fn("processfclayer")
        f.restype = None
        f(C.c_void_p(act.ctypes.data), C.c_void_p(w.ctypes.data), C.c_int32(enc), C.c_uint32(n_in),
          C.c_uint32(n_out), C.c_void_p(out.ctypes.data))
        return out

    def conv33relu(self, act: np.ndarray, w9: np.ndarray, xy: int, n_shift: int = 4) -> np.ndarray:
        act = np.ascontiguousarray(act, dtype=np.int32)
        w9 = np.ascontiguousarray(w9, dtype=np.int8)
        out = np.zeros((xy - 2) * (xy - 2), dtype=np.int32)
        f = self._fn("processconv33ReLU")
        f.restype = C.c_void_p
        f(C.c_void_p(act.ctypes.data), C.c_void_p(w9.ctypes.data), C.c_uint32(xy), C.c_uint32(n_shift),
          C.c_void_p(out.ctypes.data))
        return out

    def maxpool22(self, act: np.ndarray, xy: int) -> np.ndarray:
        act = np.ascontiguousarray(act, dtype=np.int32)
        out = np.zeros((xy // 2) * (xy // 2), dtype=np.int32)
        f = self._fn("processmaxpool22")
        f.restype = C.c_void_p
        f(C.c_void_p(act.ctypes.data), C.c_uint32(xy), C.c_void_p(out.ctypes.data))
        return out


class Oracle(_Base):
    """Our restatement (bitnet_oracle.c)."""
    prefix = "orc_"

    def __init__(self, path: Optional[str] = None):
        super().__init__(path or os.path.join(HERE, "libbnm_oracle.so"))

    def infer(self, model, images: np.ndarray, threads: int = 0, nf4_extension: bool = False):
        imgs = _as_images(images, model.img_bytes)
        n = imgs.shape[0]
        logits = np.zeros((n, model.n_classes), dtype=np.int32)
        labels = np.zeros(n, dtype=np.uint32)
        table, keep = _layer_table(model)
        f = self.lib.orc_infer_batch
        f.restype = C.c_int
        rc = f(C.c_int(model.model_class), table, C.c_uint32(len(model.layers)), C.c_void_p(imgs.ctypes.data),
               C.c_size_t(n), C.c_uint32(model.img_bytes), C.c_void_p(logits.ctypes.data),
               C.c_void_p(labels.ctypes.data), C.c_int(threads), C.c_int(int(nf4_extension)))
        if rc != 0:
            raise RuntimeError(f"orc_infer_batch failed rc={rc}")
        del keep
        return logits, labels

    def decode_fc(self, weights: np.ndarray, enc: int, n_in: int, n_out: int, nf4_extension: bool = False) -> np.ndarray:
        w = np.ascontiguousarray(weights)
        dense = np.zeros((n_out, n_in), dtype=np.int16)
        self.lib.orc_decode_fc.restype = C.c_int
        self.lib.orc_decode_fc(C.c_void_p(w.ctypes.data), C.c_int32(enc), C.c_uint32(n_in), C.c_uint32(n_out),
                               C.c_void_p(dense.ctypes.data), C.c_int(int(nf4_extension)))
        return dense

    def xorshift_images(self, n: int, img_bytes: int = 256, seed: int = 12345) -> np.ndarray:
        out = np.zeros((n, img_bytes), dtype=np.int8)
        self.lib.orc_xorshift_fill(C.c_void_p(out.ctypes.data), C.c_size_t(out.size), C.c_uint32(seed))
        return out

    def num_threads(self) -> int:
        self.lib.orc_num_threads.restype = C.c_int
        return int(self.lib.orc_num_threads())


class Reference(_Base):
    """The unmodified reference kernels compiled from /root/reference (oracle/_ref)."""
    prefix = ""

    def __init__(self, path: Optional[str] = None):
        super().__init__(path or os.path.join(HERE, "_ref", "libbitnetmcu_ref.so"))

    @staticmethod
    def available() -> bool:
        return os.path.exists(os.path.join(HERE, "_ref", "libbitnetmcu_ref.so"))

    def infer(self, model, images: np.ndarray, threads: int = 0):
        imgs = _as_images(images, model.img_bytes)
        n = imgs.shape[0]
        logits = np.zeros((n, model.n_classes), dtype=np.int32)
        labels = np.zeros(n, dtype=np.uint32)
        table, keep = _layer_table(model)
        f = self.lib.ref_infer_batch
        f.restype = C.c_int
        f(C.c_int(model.model_class), table, C.c_uint32(len(model.layers)), C.c_void_p(imgs.ctypes.data),
          C.c_size_t(n), C.c_uint32(model.img_bytes), C.c_uint32(model.n_classes), C.c_void_p(logits.ctypes.data),
          C.c_void_p(labels.ctypes.data), C.c_int(threads))
        del keep
        return logits, labels

    def num_threads(self) -> int:
        self.lib.ref_num_threads.restype = C.c_int
        return int(self.lib.ref_num_threads())


def reference_dll_labels(which: str, images: np.ndarray) -> np.ndarray:
    """Labels from the reference DLL itself (``Inference()``, BitNetMCU_MNIST_dll.c:24) for 'fc' or 'cnn'."""
    lib = C.CDLL(os.path.join(HERE, "_ref", f"Bitnet_inf_{which}.so"))
    lib.Inference.restype = C.c_uint32
    imgs = _as_images(images, 256)
    return np.array([lib.Inference(C.c_void_p(imgs[i].ctypes.data)) for i in range(imgs.shape[0])], dtype=np.uint32)
