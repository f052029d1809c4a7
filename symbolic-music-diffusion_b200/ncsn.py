"""Score networks, same names and keyword surface as the reference's models/ncsn.py.

``getattr(ncsn, FLAGS.architecture).partial(**model_kwargs)`` (train_ncsn.py:194-195) works unchanged:
  TransformerDDPM   models/ncsn.py:138-179
  DenseDDPM         models/ncsn.py:122-135 (accepts-and-ignores num_heads / num_mlp_layers)
  TransformerDDPM4  named by configs/ddpm-multi-32seq-512.cfg:1 but absent upstream; alias of TransformerDDPM
                    with the flag defaults (SURVEY D6 -- recorded assumption).
The forward pass itself is hand-written CUDA behind include/smd.h (smd_forward).
"""
from .nn import ModuleSpec

TransformerDDPM = ModuleSpec("TransformerDDPM")
TransformerDDPM4 = ModuleSpec("TransformerDDPM4")
DenseDDPM = ModuleSpec("DenseDDPM")
# models/ncsn.py:83-98.  Upstream's apply() reads an undefined `t` (broken as released); here it is the
# evident intent -- the DenseDDPM stack conditioned on `sigmas`, output divided by sigma (a score network).
DenseNCSN = ModuleSpec("DenseNCSN")
