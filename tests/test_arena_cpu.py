"""Host logic of the gradient arena (functional.GradArena): adjacent 16-byte-aligned slices, one zeroing per step up to
the high-water mark, fall-back to a private buffer when the arena is absent, full or on another device."""
import torch


def test_flat_grads_carves_adjacent_zeroed_slices_from_the_arena():
    from vit_torch_b200 import functional as Fn
    dev = torch.device("cpu")
    a, b, c = torch.nn.Parameter(torch.ones(5, 3)), torch.nn.Parameter(torch.ones(7)), torch.nn.Parameter(torch.ones(2, 2))
    arena = Fn.GradArena(64, dev)
    old = Fn.grad_arena
    try:
        Fn.grad_arena = arena
        arena.begin()
        buf1, v1 = Fn._flat_grads([a, None, b], [True, True, True], dev)
        buf2, v2 = Fn._flat_grads([c], [True], dev)
        assert v1[1] is None and v1[0].shape == (5, 3) and v1[2].shape == (7,)
        base = arena.buf.data_ptr()
        assert buf1.data_ptr() == base and buf2.data_ptr() == base + buf1.numel() * 4       # adjacent
        assert all(v.data_ptr() % 16 == 0 for v in (v1[0], v1[2], v2[0]))                   # 16-byte aligned views
        assert arena.used().numel() == buf1.numel() + buf2.numel()
        v1[0].fill_(3.0); v2[0].fill_(4.0)
        arena.begin()                                                                       # next step: zeroed, rewound
        assert arena.off == 0 and float(arena.buf.abs().sum()) == 0.0
        # a request that does not fit gets its own zero buffer instead
        big, _ = Fn._flat_grads([torch.nn.Parameter(torch.ones(100))], [True], dev)
        assert big.numel() >= 100 and big.untyped_storage().data_ptr() != arena.buf.untyped_storage().data_ptr()
        assert float(big.abs().sum()) == 0.0
    finally:
        Fn.grad_arena = old
    # without an arena every block allocates its own buffer
    buf3, _ = Fn._flat_grads([a], [True], dev)
    assert buf3.untyped_storage().data_ptr() != arena.buf.untyped_storage().data_ptr()


def test_arena_takes_its_buffer_from_an_allocator_hook():
    """A data-parallel reducer may place the arena in memory registered with its communicator
    (dist.GradAllReducer.alloc_arena); a hook that declines (None) leaves the plain allocation."""
    from vit_torch_b200 import functional as Fn
    dev = torch.device("cpu")
    mine = torch.zeros(128)
    calls = []

    def alloc(n, d):
        calls.append((n, d))
        return mine[:n]

    arena = Fn.GradArena(100, dev, alloc=alloc)
    assert calls == [(100, dev)] and arena.buf.data_ptr() == mine.data_ptr()
    assert Fn.GradArena(100, dev, alloc=lambda n, d: None).buf.numel() == 100


def test_default_data_parallel_mode():
    from vit_torch_b200 import dist as vdist
    assert [vdist.default_mode(n) for n in (1, 2, 4, 8)] == ["none", "split", "deferred", "deferred"]
    # without a process group the reducer never asks NCCL for memory
    m = torch.nn.Linear(4, 4)
    red = vdist.GradAllReducer(m, overlap=False, register_arena=True)
    assert red.alloc_arena(64, torch.device("cpu")) is None and not red.arena_registered
