import sys; sys.path.insert(0,'.')
import torch
from tests.conftest import load_gdb13, pretrained_path
from tests.test_gpu_parity import _build
from graphinvent_b200._lib import lib
from oracle import mpnn_oracle as O
fx=load_gdb13(); sd=torch.load(pretrained_path(),map_location='cpu',weights_only=False)
net=_build(O.make_constants('GGNN'), sd)
bonded = fx['edges'].sum((1,2,3))>0
for mode,name in ((0,'simt'),(1,'tc rna'),(5,'tc trunc')):
    lib.gib_set_tensor_cores(1 if mode else 0); lib.gib_tc_debug(4 if mode==5 else 0)
    with torch.no_grad(): out=net(fx['nodes'].cuda(), fx['edges'].cuda()).cpu()
    err=(out-fx['logits']).abs().max(1).values
    print(name, 'bonded max %.3e mean %.3e  bondless max %.3e' % (err[bonded].max(), err[bonded].mean(), err[~bonded].max()))
lib.gib_set_tensor_cores(1); lib.gib_tc_debug(0)
