"""Developer tool: rotated-NMS us/image at 10 000 boxes (bench.py's side metric) on its own."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from mydetection_b200 import _lib
if os.environ.get('MYDET_LIB'):
    _lib.LIB_PATH = os.environ['MYDET_LIB']        # A/B runs against a variant build

for ch in ([int(a) for a in sys.argv[1:]] or [None]):
    r = bench.rotated_nms_metric(torch.device('cuda', 0), chunks=ch)
    r['chunks'] = ch
    print(json.dumps(r))
