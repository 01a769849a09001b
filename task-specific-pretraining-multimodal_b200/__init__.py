"""mml_b200 -- B200 (sm_100a) native late-fusion multimodal training step behind MML_Suite's model surface.

Sub-modules (imported lazily so that ``import mml_b200`` works on a CPU-only box for host-side logic and tests):
  _lib      ctypes binding of libmml_b200.so (C ABI in include/mml_b200.h); raises if the library is missing
  ops       torch-tensor wrappers over the C ABI
  resnet    ResNetEncoder / ResNet18 / ResNet34 (same ctor, state_dict and forward contract as the reference)
  avmnist   AVMNIST late-fusion model: forward / train_step / validation_step / get_embeddings
  engine    the fused training / inference step (static buffers, kernel schedule, CUDA graph)
  convblock ConvBlock / MNISTAudio / MNISTImage encoders (train_avmnist.yaml) and their fused step
  mmimdb    MMIMDb gated late-fusion model (config 3) + gated_engine, its fused step
  mono      MonomodalEncoder (encoder pre-training: one ResNet encoder + Linear + CE)
  data      missing-modality patterns, device mask table, luminance table, device prefetcher
  datasets  AVMNIST / MMIMDb / MOSI dataset classes (the reference's item contract + whole pinned batches for the fused steps)
  dist      one-process-per-GPU data parallelism (NCCL) with bucketed gradient allreduce
  fedavg    FedAvg weighted aggregation
  shim      registration under the reference's YAML tags / model resolver
"""
import os as _os

# The fused step runs on 3-5 CUDA streams inside one graph and the prefetcher copies on another.  With the default of 8
# hardware work queues, unrelated streams can share a queue and serialise (measured: a background H2D copy stretched the
# 3.3 ms AVMNIST step to 4.4 ms; with 32 queues 3.5 ms).  Only effective if set before the CUDA context is created.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

__version__ = "0.1.0"
