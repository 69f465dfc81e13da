"""A CPU stand-in for cake_b200.model.Context, for tests only: the same attributes and methods, but the library behind
capi.lib() is the oracle-backed emulation of tests/fake_b200 and "device" tensors are host tensors.  It lets the Python
host logic (B200Transformer / TextModelBase / Master / WireRemote) run end to end without a GPU; the product's Context
still refuses to exist without CUDA."""
import ctypes
import weakref

import torch

from cake_b200 import capi
from cake_b200.config import CConfig
from cake_b200.model import Cache
from cake_b200.synth import TORCH_DTYPES


def use_emulation(monkeypatch, tmp_path) -> None:
    from tests.fake_b200.make_fake import build as build_fake
    so = build_fake(str(tmp_path), oracle=True)
    monkeypatch.setattr(capi, "SO_PATH", so)
    monkeypatch.setattr(capi, "_lib", None)


class CpuContext:
    def __init__(self, config, var_builder, dtype="bf16", device=0, max_seq=None, topology=None):
        self.config, self.var_builder, self.dtype, self.device = config, var_builder, dtype, device
        self.max_seq = max_seq or config.max_seq_len
        self.topology = topology or {}
        self.ccfg = CConfig.from_config(config, dtype, self.max_seq)
        self.h = ctypes.c_void_p()
        self._children = []
        capi.check(capi.lib().cake_b200_ctx_create(device, ctypes.byref(self.ccfg), ctypes.byref(self.h)))
        self.torch_dtype = TORCH_DTYPES[dtype]
        self.torch_stream = None
        self.cache = Cache(self)

    def sync(self):
        capi.check(capi.lib().cake_b200_sync(self.h))

    def launch_count(self):
        return 0

    def empty(self, *shape):
        return torch.empty(shape, dtype=self.torch_dtype)

    def to_device(self, t):
        return t.to(self.torch_dtype).contiguous().clone()

    def close(self):
        if self.h:
            for ref in self._children:
                obj = ref()
                if obj is not None:
                    obj.close()
            self._children = []
            self.cache = None
            capi.lib().cake_b200_ctx_destroy(self.h)
            self.h = None
