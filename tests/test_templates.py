import os

import numpy as np
import pytest

import cilqr_b200 as cb

REF_CFG = "/root/reference/config"
FILES = {"two_straight": "scenario_two_straight.yaml", "two_borrow": "scenario_two_borrow.yaml",
         "three_straight": "scenario_three_straight.yaml", "three_bend": "scenario_three_bend.yaml"}


@pytest.mark.skipif(not os.path.isdir(REF_CFG), reason="reference tree not mounted")
@pytest.mark.parametrize("name", cb.templates.TEMPLATE_ORDER)
def test_template_table_equals_reference_yaml(name):
    m = cb.templates.load_yaml(os.path.join(REF_CFG, FILES[name]))
    t = cb.templates.TEMPLATES[name]
    assert set(m) == set(t)
    for k in m:
        assert m[k] == t[k], k


def test_params_and_scenario_shapes():
    for name in cb.templates.TEMPLATE_ORDER:
        scn = cb.get_scenario(name)
        p = scn.params
        assert p["solve_type"] == 0 and p["max_iter"] == 100
        assert scn.borders[0] > scn.borders[1]
        assert scn.tracks.shape[0] == len(scn.ic) - 1
        assert scn.ref.size() == len(scn.ref.yaw) <= 65535
        assert np.all(np.diff(scn.ref.longitude) > 0)
    assert cb.get_scenario("two_straight").params["reference_point"] == 0
    assert cb.get_scenario("three_bend").params["reference_point"] == 1
    assert cb.get_scenario("three_straight").params["use_last_solution"] == 1
    # oncoming vehicles (yaw0 > pi/2) drive against the lane direction with yaw + pi
    tb = cb.get_scenario("two_borrow")
    assert tb.tracks[2, 1, 0] < tb.tracks[2, 0, 0] and abs(tb.tracks[2, 0, 2] - np.pi) < 1e-9
    with pytest.raises(IndexError):
        tb.obstacles_at(tb.tracks.shape[1] - 10, 50)
