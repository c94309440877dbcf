"""Growing-radius tube around a set of trajectories and its inside test (reference:
mggan/manifold.py:9-18 construction, :60-77 compute_metric / compute_inside).  Plotting helpers of
the reference (matplotlib / shapely) are out of scope."""
import numpy as np


class Manifold:
    def __init__(self, construct_set, radius):
        """construct_set (num_samples, pred_len, 2); tube radius grows linearly to `radius`."""
        self.data = np.asarray(construct_set)
        pred_len = self.data.shape[1]
        self.radius = np.linspace(radius / pred_len, radius, pred_len, endpoint=True)

    def compute_inside(self, test_data):
        """test_data (n, pred_len, 2) -> bool (n,): at every step within radius of SOME construction sample."""
        test_data = np.asarray(test_data)
        if test_data.shape[0] == 0:
            return np.zeros((0,), dtype=bool)
        d = np.linalg.norm(self.data[None] - test_data[:, None], ord=2, axis=-1)      # (n, samples, pred_len)
        return (d < self.radius[None, None]).any(1).all(1)

    def compute_metric(self, test_data):
        return np.sum(self.compute_inside(test_data)) / len(test_data)
