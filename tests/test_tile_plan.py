"""CPU tests of the cluster kernel's plan (``b2l_tile_shape_info``: pure arithmetic inside the library, no device
call): which shapes of ``pl.loo`` on the (chain, draw, obs) layout are read in place, with what cluster size, how a
long draw axis is chunked, and that the threshold ranks expect enough candidates for the tail."""
import math

import pytest

from pyloo_b200 import engine


def plan(S, reff):
    M = engine.tail_length(S, reff)
    return M, engine.tile_shape_info(S, M)


def harmonic_count(bins, q, f):
    return f * bins * sum(1.0 / i for i in range(32 - q + 1, 33))


def test_headline_shape():
    M, p = plan(4000, 1.0)
    assert M == 190
    assert p["eligible"] == 1 and p["obs_per_tile"] == 8 and p["cluster_size"] == 8 and p["n_chunks"] == 1
    assert p["draws_per_cta"] == 500 and p["boxes_per_cta"] * p["draws_per_box"] >= 500 and p["draws_per_box"] % 8 == 0
    assert p["draws_per_box"] <= 256 and p["smem_bytes"] <= 38 * 1024          # six CTAs per SM
    assert p["tail_regs_per_lane"] == 8 and p["candidate_cap"] == 512
    assert 1 <= p["rank_tight"] <= p["rank_loose"] <= 31


@pytest.mark.parametrize("S,reff,csize", [(4000, 1.0, 8), (2000, 1.0, 4), (2000, 0.5, 8), (1000, 1.0, 4), (1000, 0.3, 8),
                                          (512, 1.0, 2), (4096, 1.0, 8), (3000, 1.0, 8)])
def test_cluster_size_follows_draws_and_tail(S, reff, csize):
    """Fewest CTAs that hold the draws at <= 512 each, with at most 1.2 tail draws per threshold bin."""
    M, p = plan(S, reff)
    assert p["eligible"] == 1 and p["cluster_size"] == csize and p["n_chunks"] == 1
    assert p["draws_per_cta"] == math.ceil(S / csize) <= 512
    assert csize == 8 or M + 1 <= 1.2 * 32 * csize


@pytest.mark.parametrize("S,chunks", [(8000, 2), (16000, 4), (6000, 2), (12000, 3), (10000, 4), (16384, 4), (5000, 2)])
def test_long_draw_axes_are_cut_into_equal_chunks(S, chunks):
    M, p = plan(S, 1.0)
    assert p["eligible"] == 1 and p["n_chunks"] == chunks and p["chunk_len"] * chunks == S and p["chunk_len"] <= 4096
    assert p["cluster_size"] == 8 and p["rank_tight"] == p["rank_loose"]        # one list
    # every chunk's threshold expects ~1.6 shares of the tail: the union stays within what the tail kernel sorts
    k = harmonic_count(256, p["rank_loose"], 0.93)
    assert 1.3 * (M + 1) / chunks <= k <= 2.0 * (M + 1) / chunks
    assert chunks * k * 1.25 <= (1024 if p["tail_regs_per_lane"] == 32 else p["candidate_cap"])


@pytest.mark.parametrize("S,reff,ok", [(4000, 0.5, 1), (4000, 0.2, 1), (4000, 0.15, 1), (4000, 0.12, 0), (510, 1.0, 0),
                                       (4001, 1.0, 0), (17000, 1.0, 0), (4100, 1.0, 1), (8200, 1.0, 1), (8202, 1.0, 1), (8194, 1.0, 0),
                                       (16000, 0.5, 1), (16000, 0.3, 0)])
def test_eligibility_limits(S, reff, ok):
    """Even S from 512 to 16 384 (above 4096 divisible into <= 4 chunks of <= 4096), tails up to 510 draws (576
    chunked); everything else takes the panel route."""
    M, p = plan(S, reff)
    assert p["eligible"] == ok, (S, reff, M, p)


@pytest.mark.parametrize("reff", [1.0, 0.7, 0.5, 0.3, 0.2, 0.15])
def test_threshold_ranks_expect_enough_candidates(reff):
    """S = 4000: the loose list is expected to hold the tail with room to spare and to fit the scratch row; the tight
    list sits between M + 1 and one sort of the tail kernel."""
    M, p = plan(4000, reff)
    assert p["eligible"] == 1
    bins = 32 * p["cluster_size"]
    if M + 1 <= 0.9 * bins:
        count = lambda q: -0.9 * bins * math.log(1 - q / 32)
        k_t, k_l = count(p["rank_tight"]), count(p["rank_loose"])
    else:
        k_t, k_l = harmonic_count(bins, p["rank_tight"], 0.91), harmonic_count(bins, p["rank_loose"], 0.93)
    assert k_l >= 1.2 * (M + 1) and k_l <= 0.9 * p["candidate_cap"]
    assert M + 1 <= k_t <= 1.1 * 32 * p["tail_regs_per_lane"]   # (beyond one sort the tail kernel takes both lists)
