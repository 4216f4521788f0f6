"""GPU tests of the step engine's host<->device plumbing."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_batch_prefetcher_double_buffers_keep_batches_intact():
    """Six distinct pinned host batches through the two alternating device buffer sets, with a consumer that is slow
    enough for the next copy to be issued while the previous batch is still being read."""
    from cervix_b200.engine import BatchPrefetcher
    torch.manual_seed(0)
    host = [(torch.full((4, 3, 64, 64), float(i)).pin_memory(), torch.full((4, 64, 64), i, dtype=torch.int64).pin_memory(), None)
            for i in range(6)]
    big = torch.randn(4096, 4096, device="cuda")
    sums, seen = [], 0
    for i, (a, b, c) in enumerate(BatchPrefetcher(iter(host))):
        assert c is None and a.is_cuda and b.is_cuda
        for _ in range(3):
            big = (big @ big).clamp_(-1, 1)          # keeps the compute stream busy past the next prefetch
        sums.append((a.sum() + 0 * big[0, 0], b.sum()))
        seen += 1
    assert seen == 6
    for i, (sa, sb) in enumerate(sums):
        assert float(sa) == i * 4 * 3 * 64 * 64 and int(sb) == i * 4 * 64 * 64
