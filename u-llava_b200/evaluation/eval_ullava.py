"""validate() of /root/reference/evaluation/eval_ullava.py:33-102 on the B200 path: same arguments and return value
(ciou, giou, prec@0.5), any batch size (the reference is fixed at 1), metrics accumulated on the device.

num_workers defaults to 0 (the reference uses 4 forked workers): this build's CLIPProcessor / SegToolBox run on the
GPU, and CUDA cannot be used in a forked DataLoader worker.  Image decoding is the only per-item host work left; pass
num_workers > 0 together with a dataset whose processors stay on the host, or a 'spawn' multiprocessing context."""
import torch

from evaluation.tools import SegMeter, dict_to_cuda


def validate(model, val_dataset, data_collator, dtype, batch_size: int = 1, num_workers: int = 0, verbose: bool = True):
    model.eval()
    loader = torch.utils.data.DataLoader(val_dataset, batch_size=batch_size, shuffle=False, num_workers=num_workers,
                                         pin_memory=True, collate_fn=data_collator)
    meter = None
    for input_dict in loader:
        input_dict['inference'] = True
        input_dict = dict_to_cuda(input_dict, dtype)
        with torch.no_grad():
            out = model(**input_dict)
        if meter is None:
            meter = SegMeter(out["pred_masks"][0].device)
        meter.update(out["pred_masks"], out["gt_masks"], out["pred_boxes"], out["gt_boxes"])
    if meter is None:
        return 0.0, 0.0, 0.0
    meter.all_reduce()
    r = meter.result()
    if verbose:
        print("ciou: {:.2f}, giou: {:.2f}, prec@0.5: {:.2f} success: {}".format(r["ciou"], r["giou"], r["prec05"],
                                                                                r["images"]))
    return r["ciou"], r["giou"], r["prec05"]
