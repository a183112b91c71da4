"""`de.keras` -- the reference's import path of the embedding layers (`tfra.dynamic_embedding.keras.layers.Embedding`,
python/keras/layers/__init__.py): the torch-module mirrors of `dynamic_embedding/layers.py` under the same names,
`HvdAllToAllEmbedding` included (here the all-to-all runs over NCCL / NVLink peer memory, not Horovod).
Keras callbacks / models and the layer normalisation of python/keras/ are off the hot path."""
from . import layers  # noqa: F401
