"""CPU suite: host logic of the context model -- inverse hash tables, stream chunking, python hash twin."""
import numpy as np
import torch

from conftest import R3


def test_python_hash_twin_matches_oracle(oracle, golden):
    from cnc_b200.context_models import get_grid_index

    rng = np.random.default_rng(0)
    for D, res, T in ((3, 18, 5832), (3, 148, 2 ** 19), (3, 514, 2 ** 19), (2, 1026, 2 ** 17), (2, 130, 16904), (3, 33, 35944)):
        pos = rng.integers(0, res, (5000, D))
        got = get_grid_index(T, res, torch.from_numpy(pos)).numpy()
        want = oracle.grid_rows(pos.astype(np.uint32), T, res)
        np.testing.assert_array_equal(got, want.astype(np.int64))


def test_inverse_tables_small_cpu():
    from cnc_b200.context_models import CNC_context_models, get_grid_index

    torch.manual_seed(0)
    res3, log2T = [6, 10, 18, 34], 12
    cm = CNC_context_models(num_dim=3, resolutions_list=res3, resolutions_list_2D=[18, 34], log2_hashmap_size=log2T,
                            log2_hashmap_size_2D=10, n_features=8, sample_num=500, ste_binary=True, Rb=16,
                            skip_levels_3D=(0, 1, 2), device="cpu")
    assert cm.offs == [0, 216, 1216, 5312, 9408]
    assert cm.resolution_thresh == 10.0 and cm.n_levels_thresh == 2
    for n, r in enumerate(res3):
        T = cm.offs[n + 1] - cm.offs[n]
        uv, cnt, cs, pts = cm.unique_value_list[n], cm.unique_count_list[n], cm.unique_count_cumsum_list[n], cm.pos_grid_sorted_list[n]
        assert int(cnt.sum()) == r ** 3 == pts.shape[0] and int(cs[-1]) == r ** 3 and cs[0] == 0
        assert uv.unique().numel() == uv.numel()
        # every voxel of a group hashes to the group's row
        rows = get_grid_index(T, r, pts.to(torch.int64))
        assert torch.equal(rows, torch.repeat_interleave(uv, cnt))
        if r ** 3 <= 2 ** log2T:
            assert (cnt == 1).all() and uv.numel() == r ** 3      # dense: shuffled symbol order, one voxel per row
            assert not torch.equal(uv, uv.sort().values)
        else:
            assert torch.equal(uv, uv.sort().values) and cnt.max() > 1  # hashed: ascending row order, collisions
    assert cm.ttl_hashparams_num_valid_levels == int(cm.hashparams_num_levels[3])
    assert len(cm._chunks(3)) == 1 and cm._chunks(3)[0] == (0, int(cm.hashparams_num_levels[3]))
    assert len(cm.context_model_2D) == 1 and cm.context_model_2D[0][0].in_features == 8 * 2 + 1
    assert cm.context_model_3D[0].in_features == 25


def test_stream_chunking_matches_reference_numbers():
    """at the product config the chunk size is min(floor(2e7 / (res^3 / entries)), entries) with the ratio
    evaluated in float32 like the reference (int64 / int64 tensor division) -> 504198 / 197258 / 77216 rows."""
    for res, entries, want in ((275, 524288, 504198), (376, 524288, 197258), (514, 524288, 77216), (201, 524288, 524288)):
        ratio = ((torch.tensor(res) ** 3) / torch.tensor(entries)).item()
        per = min(int(20000000 // ratio), entries)
        assert per == want
        assert -(-entries // per) == {504198: 2, 197258: 3, 77216: 7, 524288: 1}[want]


def _small(**kw):
    from cnc_b200.context_models import CNC_context_models

    return CNC_context_models(num_dim=3, resolutions_list=[6, 10, 14, 18, 34], resolutions_list_2D=[18, 34], log2_hashmap_size=13,
                              log2_hashmap_size_2D=10, n_features=8, sample_num=500, ste_binary=True, Rb=16,
                              skip_levels_3D=(0, 1), device="cpu", **kw)


def test_dense_level_symbol_order_follows_the_reference_draw_order():
    """utils_bpp_acc.py:301,311-315: the constructor walks the levels finest first and draws one `torch.randperm(entries)`
    from the global CPU generator for every dense level (coded or skipped) -> same seed, same permutations."""
    torch.manual_seed(123)
    cm = _small()
    torch.manual_seed(123)
    want = {}
    for i in reversed(range(5)):                  # the reference's loop order
        r = cm.res[i]
        if r <= cm.resolution_thresh:
            want[i] = torch.randperm(min(r ** 3, 2 ** 13))
    assert sorted(want) == [0, 1, 2, 3]
    for i, perm in want.items():
        T = cm.offs[i + 1] - cm.offs[i]
        assert torch.equal(cm.unique_value_list[i], perm)   # dense level: row r is entry r before the shuffle
        p = cm.pos_grid_sorted_list[i].to(torch.int64)
        assert torch.equal(p[:, 0] + p[:, 1] * cm.res[i] + p[:, 2] * cm.res[i] ** 2, perm) and T >= perm.numel()
    # a different global state gives a different order (the order is not a constant)
    torch.manual_seed(124)
    assert not torch.equal(_small().unique_value_list[3], cm.unique_value_list[3])


def test_layout_rebuilds_identical_tables_in_a_fresh_object():
    """the permutation must travel with the bitstream.  `layout()` carries either the int seed or the CPU
    generator state the reference-style constructor started from; `from_layout` rebuilds the same symbol order whatever
    the global generator looks like by then."""
    import json

    from cnc_b200.context_models import CNC_context_models

    for kw in ({}, {"shuffle_seed": 77}):
        torch.manual_seed(5)
        cm = _small(**kw)
        lay = json.loads(json.dumps(cm.layout()))            # survives the container's json header
        torch.manual_seed(999)
        torch.rand(17)
        cm2 = CNC_context_models.from_layout(lay, device="cpu", sample_num=500)
        for a, b in zip(cm.unique_value_list + cm.pos_grid_sorted_list, cm2.unique_value_list + cm2.pos_grid_sorted_list):
            assert torch.equal(a, b)
        assert cm2.skip_levels_3D == (0, 1) and cm2.res == cm.res
    assert isinstance(_small(shuffle_seed=3).layout()["shuffle_seed"], int)


def test_data_parallel_share_of_the_sampled_entries_cpu():
    """CNC_context_models.set_data_parallel: every rank's share of the sampled 3D entries (sample_num / world spread over the
    levels like the single-process sample, at least one entry and at most the whole level), and the single-process setting"""
    from cnc_b200.context_models import CNC_context_models

    torch.manual_seed(0)
    cm = CNC_context_models(num_dim=3, resolutions_list=[6, 10, 18, 34], resolutions_list_2D=[18, 34], log2_hashmap_size=12,
                            log2_hashmap_size_2D=10, n_features=8, sample_num=800, ste_binary=True, Rb=16,
                            skip_levels_3D=(0, 1, 2), device="cpu")
    full = cm.sample_num_levels.clone()
    for world in (2, 4, 8):
        cm.set_data_parallel(world - 1, world)
        snl, n_valid = cm._dp_snl
        assert cm.dp_rank == world - 1 and cm.dp_world == world
        assert (snl >= 1).all() and (snl <= cm.hashparams_num_levels).all()
        assert abs(int(snl.sum()) * world - int(full.sum())) <= world * cm.n_levels      # the ranks together: the whole sample
        coded = [n for n in range(cm.n_levels) if n not in cm.skip_levels_3D and n < cm.Pg_level]
        assert n_valid == sum(int(snl[n]) for n in coded)
    cm.set_data_parallel(0, 1)
    assert cm._dp_snl is None and cm.dp_world == 1
    # the plane terms dealt round-robin cover every coded (axis, level) exactly once over the ranks
    coded_2D = [n for n in range(cm.n_levels_2D) if not (n in cm.skip_levels_2D or n >= cm.Pg_level_2D)]
    terms = [(a, n) for a in ("xy", "xz", "yz") for n in coded_2D]
    for world in (1, 2, 3, 8):
        dealt = [{t for i, t in enumerate(terms) if i % world == r} for r in range(world)]
        assert set().union(*dealt) == set(terms) and sum(len(d) for d in dealt) == len(terms)
