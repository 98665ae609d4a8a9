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
        x_again = torch.randn(tuple(x.shape), generator=gen)
        assert torch.equal(x_again, x), "generator stream does not reproduce the stored x"
        if tuple(x.shape[2:]) == (64, 64):
            lq = torch.rand(batch, 3, hw, hw, generator=gen) * 2 - 1
        else:       # non-square latents: LQ has the latent's aspect (x 2^feature-extractor stages; realsr / tiny: x 1)
            lq = torch.rand(batch, 3, x.shape[2] * hw // 64, x.shape[3] * hw // 64, generator=gen) * 2 - 1
    mask = torch.from_numpy(g["mask"]) if "mask" in g.files else None
    return x, t, lq, mask


def golden_loop_inputs(g):
    """(y, noises[T+1]) of a loop fixture; fixtures without stored noise re-draw it exactly as oracle/make_golden.py did
    (`seeded` = [seed, batch, T]: y first, then the T + 1 noise tensors, one torch.randn call each)."""
    import torch
    y = torch.from_numpy(g["y"])
    if "noises" in g.files:
        return y, torch.from_numpy(g["noises"])
    seed, batch, T = (int(v) for v in g["seeded"])
    gen = torch.Generator().manual_seed(seed)
    y_again = torch.rand(tuple(y.shape), generator=gen) * 2 - 1
    assert torch.equal(y_again, y), "generator stream does not reproduce the stored y"
    return y, torch.stack([torch.randn(tuple(y.shape), generator=gen) for _ in range(T + 1)])
