"""CPU restatement (numpy, float64) of the reference's rotated-box tracklet state machine.  TEST INFRASTRUCTURE ONLY.

Follows utils/kalman_filter.py:77-142 (RotBBoxKalmanFilter: constant-velocity model on (cx, cy, w, h, angle), process /
measurement noise of the four size-like components scaled by w*h) and utils/structures.py:447-529 (KFTracklet: angle
kept in [0, 180), measurement angle unwrapped towards the state, score momentum 0.8, feasibility test, Gaussian
likelihood of candidate boxes).  Pinned by tests/golden/tracking.npz, which the unmodified reference generated
(tests/golden/make_golden.py: gen_tracking).
"""
import numpy as np

INITIAL_P = np.array([0.1, 0.1, 0.1, 0.1, 10, 0.1, 0.1, 0.1, 0.1, 10], dtype=np.float64)      # structures.py:458-461
Q_STD = np.array([0.049, 0.032, 0.052, 0.097, 13.62, 0.01, 0.01, 0.01, 0.01, 1], dtype=np.float64)
R_STD = np.array([0.073, 0.064, 0.124, 0.163, 24.39], dtype=np.float64)
MOMENTUM = 0.8                                                                                   # structures.py:471
XYWH_ROWS = np.array([True, True, True, True, False] * 2)                                        # kalman_filter.py:94


class Bank:
    """State of N tracklets: x (N,10), P (N,10,10), score (N), pred_count (N)."""

    def __init__(self, boxes, scores):
        boxes = np.array(boxes, dtype=np.float64).reshape(-1, 5).copy()
        boxes[:, 4] = boxes[:, 4] % 180                                   # structures.py:463
        n = boxes.shape[0]
        self.x = np.concatenate([boxes, np.zeros_like(boxes)], axis=1)    # kalman_filter.py:98
        self.P = np.zeros((n, 10, 10))
        for i in range(n):
            P = np.diag(np.square(INITIAL_P))
            P[XYWH_ROWS] *= self.x[i, 2] * self.x[i, 3]                   # :100 (rows of a diagonal matrix)
            self.P[i] = P
        self.score = np.array(scores, dtype=np.float64).copy()
        self.pred_count = np.zeros(n, dtype=np.int64)
        F = np.eye(10)
        for i in range(5):
            F[i, 5 + i] = 1
        self.F = F

    def predict(self):
        """KFTracklet.predict for every tracklet (structures.py:474-485).  Returns the predicted boxes (N,5)."""
        out = np.zeros((self.x.shape[0], 5))
        for i in range(self.x.shape[0]):
            x, P = self.x[i], self.P[i]
            Q = np.diag(np.square(Q_STD))
            Q[XYWH_ROWS] *= x[2] * x[3]                                   # kalman_filter.py:108
            x = self.F @ x + 0
            P = np.linalg.multi_dot([self.F, P, self.F.T]) + Q
            out[i] = x[:5]
            x = x.copy()
            x[4] = x[4] % 180                                             # structures.py:478
            self.x[i], self.P[i] = x, P
        self.score = np.where(self.pred_count >= 1, MOMENTUM * self.score, self.score)
        self.pred_count += 1
        return out

    def update(self, boxes, scores, has):
        """KFTracklet.update for the tracklets with has[i] (structures.py:487-503).  Returns (N,5), zeros elsewhere."""
        out = np.zeros((self.x.shape[0], 5))
        H = np.eye(5, 10)
        for i in np.nonzero(has)[0]:
            x, P = self.x[i], self.P[i]
            m = np.array(boxes[i], dtype=np.float64)
            z = m[4] % 180
            m[4] = min(z, z - 180, z + 180, key=lambda v: abs(v - x[4]))
            R = np.diag(np.square(R_STD))
            R[XYWH_ROWS[:5]] *= x[2] * x[3]
            y = m - H @ x
            S = np.linalg.multi_dot((H, P, H.T)) + R
            K = np.linalg.multi_dot((P, H.T, np.linalg.inv(S)))
            x = x + K @ y
            P = P - np.linalg.multi_dot((K, H, P))
            out[i] = x[:5]
            x[4] = x[4] % 180
            self.x[i], self.P[i] = x, P
            self.score[i] = MOMENTUM * self.score[i] + (1 - MOMENTUM) * scores[i]
            self.pred_count[i] = 0
        return out

    def likelihood(self, cand):
        """KFTracklet.likelihood of M candidate boxes under every tracklet (structures.py:519-528) -> (N,M)."""
        cand = np.asarray(cand, dtype=np.float64).reshape(-1, 5)
        out = np.zeros((self.x.shape[0], cand.shape[0]))
        for i in range(self.x.shape[0]):
            mean, cov = self.x[i, :5].reshape(1, 5), self.P[i, :5, :5]
            num = ((cand - mean) @ np.linalg.inv(cov) * (cand - mean)).sum(axis=1)
            out[i] = np.exp(-0.5 * num) / np.sqrt((2 * np.pi) ** 5 * np.linalg.det(cov))
        return out

    def feasible(self, img_hw, boxes):
        """KFTracklet.is_feasible (structures.py:505-514) on the tracklets' current boxes."""
        imh, imw = img_hw
        b = np.asarray(boxes)
        bad = (self.score < 0.1) | (b[:, :4] < 0).any(axis=1) | (b[:, 0] > imw) | (b[:, 1] > imh) | (b[:, 2] > imw) | (b[:, 3] > imh)
        return ~bad
