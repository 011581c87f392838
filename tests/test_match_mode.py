"""b200slam_set_match_mode: the tile-shape policy changes the launch, never the result."""
import numpy as np
import pytest

from conftest import bits

pytestmark = pytest.mark.gpu


def test_match_modes_give_identical_results(ctx, b200slam, oracle, synth):
    w = synth.make_workload("tiny")
    field = oracle.edt(w["occ"])
    m = ctx.new_map(*field.shape)
    try:
        m.set_geometry(w["pixel"], w["top_left"]).upload_field(field)
        ctx.scan_upload(w["scan_x"], w["scan_y"])
        om = oracle.make_map(field, w["pixel"], w["top_left"])
        for n in [(3, 3, 3), (8, 32, 32), (5, 70, 130)]:
            ores, oscores, _ = oracle.score_lattice(om, w["scan_x"], w["scan_y"], w["pose0"], w["step"], n)
            for mode in (b200slam.MATCH_THROUGHPUT, b200slam.MATCH_LATENCY):
                ctx.set_match_mode(mode)
                res, scores, _ = ctx.score_lattice(m, w["pose0"], w["step"], n, want_scores=True)
                assert np.array_equal(bits(scores), bits(oscores))
                assert res.best_index == ores.best_index and res.best_hits == ores.best_hits
        with pytest.raises(b200slam.B200SlamError):
            ctx.set_match_mode(7)
    finally:
        ctx.set_match_mode(b200slam.MATCH_LATENCY)
        m.close()
