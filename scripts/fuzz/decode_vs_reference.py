"""One-off differential fuzz (CPU, needs the reference checkout): oracle/decode.py against the UNMODIFIED YOLOLayer /
RAPiDLayer / FCOSLayer (FCOS2) decode branches -- 40 seeds, random plane shapes from 1x1 to 19x19, batch 1-3, 0 / 1 / 5 /
80 classes, logits scaled by 1, 4 and 12 (exp overflow to inf included).  Last run: 330 tensors, 0 mismatches."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for sub in ('tests/golden', '', 'tests'):
    sys.path.insert(0, os.path.join(ROOT, sub))
import make_golden as m, torch, numpy as np
m.import_reference(); torch.set_grad_enabled(False)
from models.detlayers.yolov3 import YOLOLayer
from models.detlayers.fcos2 import FCOSLayer
from models.detlayers.rapid import RAPiDLayer
from oracle import decode as od
from helpers import level_anchors, YOLO_ANCHORS, RAPID_ANCHORS
bad=0; n=0
def cmp(p, got, tag):
    global bad,n
    for k,v in zip(('bbox','class_idx','score'), got):
        n+=1
        a=p[k]; 
        same = torch.equal(a, v) or (a.dtype.is_floating_point and torch.equal(torch.nan_to_num(a,nan=-7.0), torch.nan_to_num(v,nan=-7.0)))
        if not same:
            bad+=1; print('MISMATCH',tag,k,int((a!=v).sum()))
strides=[8,16,32]
for seed in range(40):
    gen=torch.Generator().manual_seed(seed)
    nb=int(torch.randint(1,4,(1,),generator=gen)); li=seed%3; s=strides[li]
    nh,nw=int(torch.randint(1,20,(1,),generator=gen)),int(torch.randint(1,20,(1,),generator=gen))
    img_hw=(nh*s,nw*s); nc=[0,1,5,80][seed%4]
    scale=[1.0,4.0,12.0][seed%3]
    # YOLO
    store,raw=m.head_views(gen,nb,3,nh,nw,4,nc); store['nchw'].mul_(scale)
    cfg={'model.yolo.anchors':m.YOLO_ANCHORS,'model.yolo.anchor_indices':m.IDX3,'model.yolo.anchor.negative_threshold':0.7,'model.fpn.out_strides':strides,'general.num_class':nc}
    p,_=YOLOLayer(li,cfg)(raw,img_hw,None)
    cmp(p, od.decode_yolo(raw, level_anchors(YOLO_ANCHORS,li), s, nc), f'yolo seed{seed}')
    # RAPiD
    store,raw=m.head_views(gen,nb,3,nh,nw,5,nc); store['nchw'].mul_(scale)
    cfgr={'model.rapid.anchors':m.RAPID_ANCHORS,'model.rapid.anchor_indices':m.IDX3,'model.fpn.out_strides':strides,'general.num_class':nc,'model.rapid.wh_smooth_l1_beta':1,'model.angle.loss_angle':'Periodic_L1','model.angle.pred_range':360}
    p,_=RAPiDLayer(li,cfgr)(raw,img_hw,None)
    cmp(p, od.decode_rapid(raw, level_anchors(RAPID_ANCHORS,li), s, nc), f'rapid seed{seed}')
    # FCOS2 (needs classes)
    if nc>0:
        store,raw=m.head_views(gen,nb,1,nh,nw,4,nc,separate=True); store['bbox_nchw'].mul_(scale); store['cls_nchw'].mul_(scale)
        cfgf={'model.fcos.anchors':[0,64,128,256,512,100000000],'model.fpn.out_strides':[8,16,32,64,128],'general.num_class':nc,'model.fcos2.ignored_threshold':0.7,'general.pred_bbox_format':'cxcywh'}
        p,_=FCOSLayer(li,cfgf)(raw,img_hw,None)
        cmp(p, od.decode_fcos(raw, s, img_hw), f'fcos seed{seed}')
print('compared',n,'bad',bad)
