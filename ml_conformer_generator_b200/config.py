"""Constants of the hot path.  Values follow the reference's `src/mlconfgen/utils/config.py:3-32` and the
hyper-parameters hard-coded in `src/mlconfgen/conformer_generator.py:67-88`."""

DIMENSION = 42
NUM_BOND_TYPES = 5
CONTEXT_NORMS = {
    "mean": [105.0766, 473.1938, 537.4675],
    "mad": [52.0409, 219.7475, 232.9718],
}
ATOM_DECODER = {0: "C", 1: "N", 2: "O", 3: "F", 4: "P", 5: "S", 6: "Cl", 7: "Br"}
PERMITTED_ELEMENTS = (6, 7, 8, 9, 15, 16, 17, 35)
MIN_N_NODES = 15
MAX_N_NODES = 39

# conformer_generator.py:67-88
HIDDEN_NF = 420
IN_NODE_NF = 12  # 8 atom classes + time + 3 context
N_CLASSES = 8
N_BLOCKS = 9
NORMALIZATION_FACTOR = 100.0
NOISE_PRECISION = 1e-5
SEER_HIDDEN = 2048
SEER_EMBEDDING_DIM = 64
SEER_NUM_EMBEDDINGS = 36
