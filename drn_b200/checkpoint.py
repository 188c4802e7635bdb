"""Checkpoint compatibility with the reference driver (main.py:104-111 resume, 369-373 save).

main.py saves `{'epoch', 'state_dict', 'loss', 'top1', 'top5'}` where `state_dict` comes from the `nn.DataParallel` wrapper
(keys carry the `module.` prefix) and resumes with a key-matched PARTIAL load (`{k: v for k, v in pretrained.items() if k in
model_dict}`).  Because `mainModel` keeps the reference's parameter names and shapes, the published DRN checkpoints load into
it unchanged; these helpers do the same for a bare (unwrapped) model and write checkpoints the reference can resume from."""
import torch


def load_reference_checkpoint(model, checkpoint, strict_shapes=True):
    """checkpoint: the dict main.py saved (or a path to it).  Returns (loaded_keys, skipped_keys) like main.py's partial load:
    keys absent from the model are skipped; with strict_shapes a shape mismatch raises instead of being silently dropped."""
    if isinstance(checkpoint, (str, bytes)):
        checkpoint = torch.load(checkpoint, map_location="cpu")
    sd = checkpoint["state_dict"] if "state_dict" in checkpoint else checkpoint
    own = model.state_dict()
    wrapped = all(k.startswith("module.") for k in own)
    loaded, skipped = [], []
    new = {}
    for k, v in sd.items():
        kk = k
        if not wrapped and kk.startswith("module."):
            kk = kk[len("module."):]
        elif wrapped and not kk.startswith("module."):
            kk = "module." + kk
        if kk in own:
            if tuple(own[kk].shape) != tuple(v.shape):
                if strict_shapes:
                    raise RuntimeError("checkpoint tensor %s has shape %s, model expects %s" % (k, tuple(v.shape), tuple(own[kk].shape)))
                skipped.append(k)
                continue
            new[kk] = v
            loaded.append(k)
        else:
            skipped.append(k)
    own.update(new)
    model.load_state_dict(own)
    return loaded, skipped


def reference_checkpoint(model, epoch=0, loss=0.0, top1=0.0, top5=0.0):
    """The dict main.py:369-373 would save for this model (`module.`-prefixed keys, as from the DataParallel wrapper)."""
    sd = model.state_dict()
    if not all(k.startswith("module.") for k in sd):
        sd = {"module." + k: v for k, v in sd.items()}
    return {"epoch": epoch, "state_dict": sd, "loss": loss, "top1": top1, "top5": top5}
