import sys, json, os
sys.argv=['bench.py','--no-cpu-baseline']
sys.path.insert(0,'.')
import bench, torch
a=bench.parse()
ctx=bench.dist_context()
from deadtrees_b200.network.segmodel import SemSegment
net=dict(architecture="unet",encoder_name="resnet34",encoder_depth=5,encoder_weights=None,decoder_channels=[256,128,64,32,16],losses=["DICE","FOCAL"],classes=["bg","a","b"],in_channels=3,precision="bf16")
torch.manual_seed(0)
seg=SemSegment(net,dict(learning_rate=3e-4,cosineannealing_tmax=10)).eval().cuda()
eng=seg.model.engine()
for b in (16,32,16,32):
    os.environ["DT_CFG5_BATCH"]=str(b)
    o=bench.cfg5_measure(a,ctx,eng,10,3)
    print(b, round(o['ms_per_step'],3), round(o['e2e']['ms_per_step'],3), round(o['roofline']['frac'],3), flush=True)
