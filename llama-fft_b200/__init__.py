"""llama-fft_b200 -- B200-native block-circulant (BC) linear hot path of Pinguhat/llama-fft.

Host side mirrors the reference's module names so callers switch by changing one import:

    reference                         here
    src/patch_llama_fft.py      ->    llama_fft_b200.patch_llama_fft
    src/fft_utils.py            ->    llama_fft_b200.fft_utils
    src/bench_all_in_one.py     ->    llama_fft_b200.bench_all_in_one   (enable/disable_cfft_cache only)
    src/calibrate_bc.py         ->    llama_fft_b200.calibrate_bc

The forward itself is hand-written sm_100a CUDA behind the C ABI in include/bclinear.h.
"""
from .patch_llama_fft import (  # noqa: F401
    MODEL_PATH, BLOCK_SIZE, NUM_LAYERS_TO_PATCH,
    BlockCirculantLinear, patch_mlp_with_block_circulant, patch_llama_fft,
    save_bc_params, load_bc_params,
    dense_block_to_circulant_column, dense_block_to_circulant_column_loss_aware, circulant_from_first_col,
    enable_cfft_cache, disable_cfft_cache, _detect_best_convention_for_layer, _get_submodule_by_name,
)
from .fft_utils import circulant_matvec_fft  # noqa: F401
from .grouped import GroupedBCForward, HostPipe, decode_tc_spectrum  # noqa: F401
from .shared_input import SharedInputGroup, fuse_shared_inputs, unfuse_shared_inputs  # noqa: F401
from .chain import BCChain, FusedBCMLP, fuse_mlp, unfuse_mlp, barrier_counter  # noqa: F401
from .sharded import ShardedBlockCirculantLinear, shard_bc_linear, shard_rows, gather_columns, push_columns  # noqa: F401

__version__ = "0.1.0"
