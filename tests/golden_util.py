"""Helpers shared by the CPU (oracle) and GPU parity tests for reading the golden fixtures."""
def golden_inputs(g):
    """(x, t, lq, mask) of a forward fixture as torch CPU tensors.  Fixtures with a large LQ input store only the seed
    (`lq_seeded` = [seed, batch, size]); the tensors are re-drawn exactly as oracle/make_golden.py drew them."""
    import torch
    x = torch.from_numpy(g["x"])
    t = torch.from_numpy(g["t"])
    if "lq" in g.files:
        lq = torch.from_numpy(g["lq"])
    else:
        seed, batch, hw = (int(v) for v in g["lq_seeded"])
        gen = torch.Generator().manual_seed(seed)
        x_again = torch.randn(batch, x.shape[1], 64, 64, generator=gen)
        assert torch.equal(x_again, x), "generator stream does not reproduce the stored x"
        lq = torch.rand(batch, 3, hw, hw, generator=gen) * 2 - 1
    mask = torch.from_numpy(g["mask"]) if "mask" in g.files else None
    return x, t, lq, mask
