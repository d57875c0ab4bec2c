"""One training step of the benchmark configuration between cudaProfilerStart/Stop, for
  ncu --profile-from-start off --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches.csv python bench/ncu_step.py
(every launch of the step with its device time; cold-cache and serialised: compare SHARES)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from distributed_vgg_f_b200.data import transforms as T
from distributed_vgg_f_b200.data.loader import FusedBatch
from distributed_vgg_f_b200.data.synthetic import synthetic_uint8_batch
from distributed_vgg_f_b200.engine.native_engine import NativeEngine
from distributed_vgg_f_b200.models.vggf import vggf_spec

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
B = int(os.environ.get("NCU_BATCH", "64"))
eng = NativeEngine(vggf_spec(3), device=dev, batch=B, lr=1e-5, seed=0)
imgs, labels = synthetic_uint8_batch(B, 128, 3, seed=0)
batch = FusedBatch(torch.from_numpy(imgs).to(dev), T.sample_train_params(B, 128, 128).to(dev),
                   torch.from_numpy(labels).to(dev), (256, 256), None)
for _ in range(2):
    eng.train_step(batch)
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.train_step(batch)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
