"""Inference configuration dictionary (same keys and defaults as the reference
util/configuration.py:138-162; the argparse training Configuration is out of scope)."""

VIDEO_INFERENCE_CONFIG = {
    'buffer_size': 100,
    'deep_update_every': -1,
    'enable_long_term': True,
    'enable_long_term_count_usage': True,
    'fbrs_model': 'saves/fbrs.pth',
    'hidden_dim': 64,
    'images': None,
    'key_dim': 64,
    'max_long_term_elements': 10000,
    'max_mid_term_frames': 10,
    'mem_every': 10,
    'min_mid_term_frames': 5,
    'model': './saves/XMem.pth',
    'no_amp': False,
    'num_objects': 1,
    'num_prototypes': 128,
    's2m_model': 'saves/s2m.pth',
    'size': 480,
    'top_k': 30,
    'value_dim': 512,
    'masks_out_path': None,
    'workspace': None,
    'save_masks': True,
}
