"""Frame preprocessing restatement (oracle/preprocess_ref.py) vs the reference's own preprocess_observation executed in
the build container (tests/golden/preprocess_cases.npz, oracle/make_golden.py).  CPU only."""
import os

import numpy as np

from oracle import preprocess_ref

G = os.path.join(os.path.dirname(__file__), 'golden')


def test_preprocess_matches_reference_execution():
    g = np.load(os.path.join(G, 'preprocess_cases.npz'))
    for game in ('Breakout', 'SpaceInvaders', 'MsPacman'):
        out = preprocess_ref.preprocess(g[game + '_obs'], game)
        assert out.dtype == np.uint8 and out.shape == g[game + '_out'].shape
        assert np.array_equal(out, g[game + '_out']), game
    assert g['Breakout_out'][1].max() <= 255 and 6 in np.unique(g['Breakout_out'][1])  # white wraps: 262.65 -> 6


def test_every_rgb_triple_matches_the_reference():
    g = np.load(os.path.join(G, 'preprocess_cases.npz'))
    assert preprocess_ref.cube_crc() == int(g['cube_crc'])
