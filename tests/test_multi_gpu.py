"""The multi-device context of the C ABI (b200ani_create_multi): needs at least two GPUs (run under `gpurun --gpus 2`)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _n_gpus():
    try:
        import torch

        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.skipif(_n_gpus() < 2, reason="needs two GPUs")
def test_multi_context_equals_single_device(engine):
    """Genomes replicated on two devices, pairs dealt, one NCCL all-gather inside the library: the array that comes back is
    byte for byte the single-device result, in caller order; same for ANIb."""
    import pyani_b200
    from pyani_b200 import synth

    gs = synth.make_set(3, 3, 250_000, seed=31)
    pairs = synth.all_directed_pairs(len(gs))[::-1]
    base = len(engine)
    gids = [engine.add_genome(g.residues, g.contig_off, g.name) for g in gs]
    want = engine.anim(np.array([[gids[a], gids[b]] for a, b in pairs], np.int32))
    wantb = engine.anib(np.array([[gids[a], gids[b]] for a, b in pairs[:20]], np.int32))
    with pyani_b200.MultiEngine([0, 1]) as me:
        ids = [me.add_genome(g.residues, g.contig_off, g.name) for g in gs]
        assert ids == list(range(len(gs)))
        got = me.anim(pairs)
        assert got.tobytes() == want.tobytes()
        assert me.anib(pairs[:20]).tobytes() == wantb.tobytes()
        me.clear()
        assert me.add_genome(gs[0].residues, gs[0].contig_off) == 0
    assert base >= 0
