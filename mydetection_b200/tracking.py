"""Tracklets of rotated boxes, batched on the GPU.

The reference keeps one `KFTracklet` per object (utils/structures.py:447-529), each with its own numpy
`RotBBoxKalmanFilter` (utils/kalman_filter.py:77-142), and a tracker steps them in Python loops.  `TrackletBank` holds the
state of ALL tracklets of a stream on the device and advances them with one launch per operation (mydet_kf_*), with the
same arithmetic (float64), the same angle conventions and the same score bookkeeping; the association inputs -- rotated
IoU of the predicted boxes with the frame's detections, and the Gaussian likelihood of the detections under every
tracklet -- come from the same library (mydet_iou_rot_pairwise, mydet_kf_likelihood).
"""
import ctypes

import torch

from . import _lib, ops

# KFTracklet.__init__ (utils/structures.py:458-461) and its score momentum (:471)
INITIAL_P = [0.1, 0.1, 0.1, 0.1, 10, 0.1, 0.1, 0.1, 0.1, 10]
Q_STD = [0.049, 0.032, 0.052, 0.097, 13.62, 0.01, 0.01, 0.01, 0.01, 1]
R_STD = [0.073, 0.064, 0.124, 0.163, 24.39]
MOMENTUM = 0.8


def _stream():
    return torch.cuda.current_stream().cuda_stream


class TrackletBank:
    """N tracklets: x (N,10), P (N,10,10), score (N) float64 and pred_count (N) int32 on the device.

    boxes: (N,5) (cx, cy, w, h, degrees), scores: (N,).  img_hw is what KFTracklet.is_feasible tests against."""

    def __init__(self, boxes, scores, img_hw=None, device=None, initial_p=INITIAL_P, q_std=Q_STD, r_std=R_STD, momentum=MOMENTUM):
        if not torch.cuda.is_available():
            raise _lib.MydetError('mydetection_b200 needs a CUDA device (B200); there is no CPU fallback')
        dev = torch.device(device) if device is not None else torch.device('cuda', torch.cuda.current_device())
        boxes = torch.as_tensor(boxes, dtype=torch.float64).reshape(-1, 5).to(dev).contiguous()
        n = boxes.shape[0]
        self.device, self.img_hw = dev, img_hw
        self._noise = (ctypes.c_double * 26)(*initial_p, *q_std, *r_std, momentum)
        self.x = torch.empty(n, 10, dtype=torch.float64, device=dev)
        self.P = torch.empty(n, 10, 10, dtype=torch.float64, device=dev)
        self.score = torch.as_tensor(scores, dtype=torch.float64).reshape(-1).to(dev).clone()
        self.pred_count = torch.zeros(n, dtype=torch.int32, device=dev)
        self.bbox = boxes.clone()                      # KFTracklet.bbox: the last predicted / updated box
        self.bbox[:, 4] = torch.remainder(self.bbox[:, 4], 180)
        with torch.cuda.device(dev):
            rc = _lib.lib().mydet_kf_initiate(ops._ptr(boxes), n, self._noise, ops._ptr(self.x), ops._ptr(self.P),
                                              ops._ptr(self.pred_count), _stream())
        _lib.check(rc, 'mydet_kf_initiate')

    def __len__(self):
        return self.x.shape[0]

    def predict(self):
        """KFTracklet.predict (structures.py:474-485) for every tracklet -> predicted boxes (N,5) float64."""
        out = torch.empty(len(self), 5, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            rc = _lib.lib().mydet_kf_predict(ops._ptr(self.x), ops._ptr(self.P), ops._ptr(self.score), ops._ptr(self.pred_count),
                                             len(self), self._noise, ops._ptr(out), _stream())
        _lib.check(rc, 'mydet_kf_predict')
        self.bbox = out
        return out.clone()

    def update(self, boxes, scores, has=None):
        """KFTracklet.update (structures.py:487-503) for the tracklets with has[i] (default: all).  boxes (N,5), scores (N).
        Returns the updated boxes (N,5); rows without a measurement are zero and keep their predicted `bbox`."""
        if len(self) and int(self.pred_count.min()) <= 0:
            raise AssertionError('Please call predict() before update()')        # structures.py:489
        boxes = torch.as_tensor(boxes, dtype=torch.float64).reshape(-1, 5).to(self.device).contiguous()
        scores = torch.as_tensor(scores, dtype=torch.float64).reshape(-1).to(self.device).contiguous()
        if boxes.shape[0] != len(self) or scores.shape[0] != len(self):
            raise ValueError('one measurement row per tracklet (use `has` to mark the tracklets that have none)')
        has_t = None if has is None else torch.as_tensor(has).to(self.device, torch.uint8).contiguous()
        out = torch.empty(len(self), 5, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            rc = _lib.lib().mydet_kf_update(ops._ptr(self.x), ops._ptr(self.P), ops._ptr(self.score), ops._ptr(self.pred_count),
                                            ops._ptr(boxes), ops._ptr(scores), ops._ptr(has_t), len(self), self._noise,
                                            ops._ptr(out), _stream())
        _lib.check(rc, 'mydet_kf_update')
        self.bbox = out if has_t is None else torch.where(has_t.bool()[:, None], out, self.bbox)
        return out

    def likelihood(self, xywha):
        """KFTracklet.likelihood (structures.py:519-528) of M candidate boxes under every tracklet -> (N,M) float64."""
        cand = torch.as_tensor(xywha, dtype=torch.float64).to(self.device).contiguous()
        assert cand.dim() == 2 and cand.shape[1] == 5
        out = torch.empty(len(self), cand.shape[0], dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            rc = _lib.lib().mydet_kf_likelihood(ops._ptr(self.x), ops._ptr(self.P), len(self), ops._ptr(cand), cand.shape[0],
                                                ops._ptr(out), _stream())
        _lib.check(rc, 'mydet_kf_likelihood')
        return out

    def association_iou(self, det_boxes):
        """Rotated IoU (N,M) float64 of the tracklets' current boxes with a frame's detections (M,5): the association
        cost a tracker thresholds / assigns on, from the same kernel as iou_rle."""
        return ops.iou_rot(self.bbox.to(torch.float32), torch.as_tensor(det_boxes, dtype=torch.float32).to(self.device))

    def is_feasible(self):
        """KFTracklet.is_feasible (structures.py:505-514) for every tracklet -> (N,) bool."""
        imh, imw = self.img_hw
        b = self.bbox
        bad = (self.score < 0.1) | (b[:, :4] < 0).any(dim=1) | (b[:, 0] > imw) | (b[:, 1] > imh) | (b[:, 2] > imw) | (b[:, 3] > imh)
        return ~bad
