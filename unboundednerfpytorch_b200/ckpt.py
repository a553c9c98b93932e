"""Checkpoint interchange with the reference's on-disk format (plain ``torch.save`` dicts named ``*.tar``):
``{'global_step', 'model_kwargs', 'model_state_dict', 'optimizer_state_dict'}`` -- FourierGrid/utils.py:61-74,
FourierGrid_ckpt_manager.py:17-57 (``load_all_info``, ``save_model``, ``load_model``) and the block merge of :60-97.

A reference-trained ``fine_last.tar`` loads into this library's model classes (same constructor keywords and state-dict names;
grids are copied into the channels-last layout on load) and a checkpoint written here loads into the reference (grids are
written in the reference's contiguous [P,C,X,Y,Z] layout).  torch >= 2.6 defaults ``torch.load`` to ``weights_only=True`` while
reference checkpoints carry NumPy arrays in ``model_kwargs`` (xyz_min / xyz_max, FourierGrid_model.py:353-354): they are loaded
with ``weights_only=False`` -- only load checkpoints you trust, exactly as with the reference."""
import torch


def _load(ckpt_path, map_location='cpu'):
    return torch.load(ckpt_path, map_location=map_location, weights_only=False)


def load_model(model_class, ckpt_path, device=None):
    """utils.load_model / FourierGridCheckpointManager.load_model -> model (on ``device`` when given)."""
    ckpt = _load(ckpt_path)
    model = model_class(**ckpt['model_kwargs'])
    model.load_state_dict(ckpt['model_state_dict'])
    return model.to(device) if device is not None else model


def load_checkpoint(model, optimizer, ckpt_path, no_reload_optimizer):
    """utils.load_checkpoint / load_all_info -> (model, optimizer, start_step).  Optimizer moments keep the parameter layout."""
    ckpt = _load(ckpt_path, map_location=next(model.parameters()).device)
    model.load_state_dict(ckpt['model_state_dict'])
    if not no_reload_optimizer and optimizer is not None:
        optimizer.load_state_dict(ckpt['optimizer_state_dict'])
        for p, st in optimizer.state.items():             # moments saved by the reference are contiguous: match the parameter layout
            for k in ('exp_avg', 'exp_avg_sq'):
                if k in st and st[k].stride() != p.stride():
                    st[k] = torch.empty_like(p, memory_format=torch.preserve_format).copy_(st[k])
    return model, optimizer, ckpt['global_step']


def save_checkpoint(global_step, model, optimizer, save_path):
    """FourierGridCheckpointManager.save_model: same keys; tensors are written contiguous so the reference can load them."""
    sd = {k: v.detach().contiguous() for k, v in model.state_dict().items()}
    osd = optimizer.state_dict() if optimizer is not None else {}
    for st in osd.get('state', {}).values():
        for k, v in list(st.items()):
            if torch.is_tensor(v):
                st[k] = v.detach().contiguous()
    torch.save({'global_step': global_step, 'model_kwargs': model.get_kwargs(), 'model_state_dict': sd,
                'optimizer_state_dict': osd}, save_path)


@torch.no_grad()
def merge_blocks(model_class, paths, device):
    """FourierGridCheckpointManager.merge_blocks (:60-97): element-wise minimum of the block models' grids and rgbnet tensors,
    the mask cache rebuilt from the merged density (update_occupancy_cache)."""
    merged = load_model(model_class, paths[0], device)
    sd = {k: v.clone() for k, v in merged.state_dict().items()}
    for path in paths[1:]:
        cur = load_model(model_class, path, device).state_dict()
        for key in sd:
            if key in ('density.grid', 'k0.grid') or 'rgb' in key:
                sd[key] = torch.min(sd[key], cur[key])
    sd.pop('mask_cache.mask', None)
    merged.load_state_dict(sd, strict=False)
    merged.update_occupancy_cache()
    return merged
