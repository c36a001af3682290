"""The small data set of the team tests (shared by the test and its worker processes)."""
import numpy as np

import oracle_lib as ol

ELL = np.array([[0, 0, 0, 60, 40, 50, 1.0], [20, -10, 5, 20, 25, 15, 0.5], [-25, 15, -10, 15, 10, 20, -0.4],
                [5, 30, 20, 12, 18, 9, 0.8]])


def make_scene(n=11, n_u=160, n_v=128, n_a=96, n_t=96):
    Ps = ol.circular_trajectory(n, 750, 1200, n_u, n_v, 200, 2.0)
    imgs = np.stack([ol.project_ellipsoids(P, n_u, n_v, ELL) for P in Ps]).astype(np.float32)
    return dict(n=n, n_u=n_u, n_v=n_v, n_a=n_a, n_t=n_t, Ps=Ps, imgs=imgs, dkappa=float(np.deg2rad(0.05)))
