"""csl_host_gate: the host-released gate bench.py puts in front of short timed regions (include/cosmoslik_b200.h).
Work behind it must not start before the host writes the flag, must start once it does, and a gate nobody opens must
give up after its timeout instead of hanging the stream."""
import time

import pytest

from cosmoslik_b200 import _lib as L

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch():
    import torch
    return torch


def test_gate_holds_the_stream_until_the_host_opens_it(torch):
    lib = L.lib()
    flag = torch.zeros(1, dtype=torch.int32).pin_memory()
    x = torch.zeros(1 << 20, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    done = torch.cuda.Event()
    L.check(lib.csl_host_gate(flag.data_ptr(), 10.0, L.stream_ptr()), "csl_host_gate")
    x.add_(1.0)
    done.record()
    time.sleep(0.05)
    assert not done.query(), "work behind a closed gate has run"
    flag[0] = 1
    t0 = time.perf_counter()
    done.synchronize()
    assert time.perf_counter() - t0 < 1.0
    assert float(x.sum()) == float(1 << 20)


def test_gate_that_is_already_open_costs_nothing_and_a_forgotten_one_times_out(torch):
    lib = L.lib()
    flag = torch.ones(1, dtype=torch.int32).pin_memory()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    L.check(lib.csl_host_gate(flag.data_ptr(), 10.0, L.stream_ptr()), "csl_host_gate")
    b.record(); torch.cuda.synchronize()
    assert a.elapsed_time(b) < 5.0
    flag[0] = 0
    a.record()
    L.check(lib.csl_host_gate(flag.data_ptr(), 0.05, L.stream_ptr()), "csl_host_gate")
    b.record(); torch.cuda.synchronize()
    assert 40.0 < a.elapsed_time(b) < 2000.0
    with pytest.raises(RuntimeError):
        L.check(lib.csl_host_gate(flag.data_ptr(), 1e6, L.stream_ptr()), "csl_host_gate")      # timeouts are capped at 30 s
    with pytest.raises(RuntimeError):
        L.check(lib.csl_host_gate(None, 1.0, L.stream_ptr()), "csl_host_gate")
