"""distributed-vgg-f_b200: a Blackwell-native data-parallel VGG-F trainer.

Public surface (mirrors the reference's two modules, distributedVggf.py / distributedUtil.py):
    vgg_funnel_model, DataManager, Trainer, manage_training, parse_command_line,
    distributed_is_initialized, Average, Accuracy2
plus the B200-native pieces: engine.NativeEngine, ops (sm_100a kernels), parallel (symmetric
arena, fused all-reduce, bucket plan).
"""
from .config import DATA, TRAIN
from .models.vggf import get_spec, vgg16_spec, vgg_funnel_model, vggf_spec
from .parallel.process_group import distributed_is_initialized
from .utils.metrics import Accuracy2, Average

__version__ = "0.1.0"


def __getattr__(name):      # lazy: these pull in heavier modules
    if name == "DataManager":
        from .data.loader import DataManager
        return DataManager
    if name == "Trainer":
        from .trainer import Trainer
        return Trainer
    if name == "manage_training":
        from .train import manage_training
        return manage_training
    if name == "parse_command_line":
        from .cli import parse_command_line
        return parse_command_line
    if name == "NativeEngine":
        from .engine.native_engine import NativeEngine
        return NativeEngine
    raise AttributeError(name)
