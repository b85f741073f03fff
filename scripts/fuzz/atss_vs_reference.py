"""One-off differential fuzz (CPU, needs the reference checkout): oracle/atss.py against the UNMODIFIED
FCOS_ATSS_Layer.forward(raw, img_size, labels) -- 25 seeds x 5 levels, 4 image shapes, 0-60 GT per image (incl. images
without GT), every third seed with GT centres snapped to cell boundaries (equidistant anchors).
Last run: 625 target tensors compared, 0 mismatches."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for sub in ('tests/golden', ''):
    sys.path.insert(0, os.path.join(ROOT, sub))
import make_golden as m, torch, numpy as np
m.import_reference(); torch.set_grad_enabled(False)
from models.detlayers.fcos2 import FCOS_ATSS_Layer
from utils.structures import ImageObjects
from oracle import atss as oa
strides=[8,16,32,64,128]
bad=0; total=0
for seed in range(25):
    gen=torch.Generator().manual_seed(5000+seed)
    img_hw = [(384,512),(640,640),(512,384),(384,384)][seed%4]
    n_cls = 6
    cfg = {'model.fpn.out_strides': strides, 'general.num_class': n_cls, 'model.fcos2.ignored_threshold': 0.7,
           'model.atss.anchors': [24,48,96,192,384], 'model.atss.topk_per_level': 9}
    labels=[]; gts=[]
    for b in range(2):
        n = int(torch.randint(0, 40, (1,), generator=gen)) if b else int(torch.randint(1, 60, (1,), generator=gen))
        bx=torch.empty(n,4)
        bx[:,0]=torch.rand(n,generator=gen)*(img_hw[1]-10)+5; bx[:,1]=torch.rand(n,generator=gen)*(img_hw[0]-10)+5
        if seed%3==0: bx[:,:2] = torch.round(bx[:,:2]/8)*8     # centres on cell boundaries: equidistant anchors
        bx[:,2:4]=torch.exp(torch.rand(n,2,generator=gen)*np.log(40)+np.log(8))
        ct=torch.randint(0,n_cls,(n,),generator=gen)
        labels.append(ImageObjects(bx,ct,bb_format='cxcywh',img_hw=img_hw)); gts.append((bx,ct))
    for li,s in enumerate(strides):
        store, raw = m.head_views(gen, 2, 1, img_hw[0]//s, img_hw[1]//s, 4, n_cls, separate=True)
        layer = FCOS_ATSS_Layer(li,cfg); grabbed={}
        def prof(frame,event,arg):
            if event=='return' and frame.f_code.co_name=='forward' and 'PositiveMask' in frame.f_locals:
                for k in ('PositiveMask','IgnoredMask','TargetConf','TargetLTRB','TargetCls'): grabbed[k]=frame.f_locals[k].clone()
        sys.setprofile(prof)
        try: layer(raw,img_hw,labels)
        except Exception as e:
            sys.setprofile(None); print('reference raised', seed, li, type(e).__name__, e); continue
        finally: sys.setprofile(None)
        t = store['bbox_nchw'].permute(0,2,3,1)
        out = oa.assign_level(li,t,gts,img_hw,strides,[24,48,96,192,384],9,0.7,n_cls)
        for k,v in out.items():
            total+=1
            if not torch.equal(v, grabbed[k]):
                bad+=1; d=(v!=grabbed[k]); print('MISMATCH seed',seed,'level',li,k,'cells',int(d.sum()), 'boundary-centres' if seed%3==0 else '')
print('compared',total,'bad',bad)
