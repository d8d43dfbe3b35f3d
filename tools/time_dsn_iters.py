import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, warnings, contextlib, io
from dasr_b200.dsn.loss import GeneratorLoss
from dasr_b200.dsn.model import De_resnet, Discriminator
from dasr_b200.dsn.train import train_iteration
from oracle import srn_oracle as O
dev = torch.device('cuda', 0)
with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
    warnings.simplefilter('ignore')
    mg = De_resnet(8, 4).to(dev); md = Discriminator(kernel_size=5, D_arch='FSD', norm_layer='Instance', filter_type='wavelet', cs='cat').to(dev)
    gl = GeneratorLoss(per_type='VGG', filter='wavelet', kernel_size=5, w_col=1, w_tex=0.005, w_per=0.01, wgan=False).to(dev)
og = torch.optim.Adam(mg.parameters(), lr=1e-4, betas=[0.5, 0.999]); od = torch.optim.Adam(md.parameters(), lr=1e-4, betas=[0.5, 0.999])
inp = O.synth_image((8, 3, 256, 256), 700).to(dev); bic = O.synth_image((8, 3, 64, 64), 800).to(dev); dis = O.synth_image((8, 3, 64, 64), 900).to(dev)
for i in range(10):
    torch.cuda.synchronize(); t0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    train_iteration(mg, md, gl, og, od, inp, bic, dis, log=False)
    e1.record(); torch.cuda.synchronize()
    print('iter %d: wall %.1f ms  events %.1f ms  mem %.1f GB reserved %.1f GB' % (i, (time.time() - t0) * 1e3, e0.elapsed_time(e1),
          torch.cuda.memory_allocated() / 2**30, torch.cuda.memory_reserved() / 2**30), flush=True)
