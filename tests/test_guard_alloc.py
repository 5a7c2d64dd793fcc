"""The guard-byte allocator behind tests/test_gpu_guards.py detects writes on either side of a buffer (CPU tensors)."""
import pytest
import torch


def test_guarded_alloc_detects_overrun(monkeypatch):
    import jampr_plus_b200.routing.env as envmod
    monkeypatch.setattr(envmod, "GUARD_BYTES", 100)           # rounded up to 256
    a = envmod._GuardedAlloc()
    t = a.zeros((3, 5), torch.int16, "cpu", "probe")
    assert t.shape == (3, 5) and t.dtype == torch.int16 and int(t.abs().sum()) == 0
    assert t.data_ptr() % 256 == a.fenced[0][1].data_ptr() % 256      # payload keeps the allocation's alignment
    t.fill_(7)
    assert a.check() == 1
    label, raw, n, g = a.fenced[0]
    assert (n, g) == (30, 256)
    raw[g + n] = 0                                             # first byte after the payload
    with pytest.raises(RuntimeError, match="probe"):
        a.check()
    raw[g + n] = 0xA5
    raw[g - 1] = 1                                             # last byte before the payload
    with pytest.raises(RuntimeError, match="probe"):
        a.check()


def test_guards_off_by_default():
    import jampr_plus_b200.routing.env as envmod
    assert envmod.GUARD_BYTES == 0
    a = envmod._GuardedAlloc()
    t = a.zeros((4,), torch.float32, "cpu")
    assert a.fenced == [] and a.check() == 0 and t.numel() == 4
