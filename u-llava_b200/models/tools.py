"""Generation helpers kept for API parity with /root/reference/models/tools.py (stopping-criteria
interface used by inference_ullava.py:88-102).  The tokenizer/embedding resize helpers of the reference
are training-time utilities and are out of scope of the inference hot path."""
import torch


class KeywordsStoppingCriteria:
    """Stops when one of `keywords` appears in the decoded continuation of sample 0 (batch-1 semantics,
    reference models/tools.py:11-31)."""

    def __init__(self, keywords, tokenizer, input_ids):
        self.keywords = keywords
        ids = [tokenizer(k).input_ids for k in keywords]
        self.keyword_ids = [i[0] for i in ids if type(i) is list and len(i) == 1]
        self.tokenizer = tokenizer
        self.start_len = None
        self.input_ids = input_ids

    def __call__(self, output_ids: torch.LongTensor, scores: torch.FloatTensor, **kwargs) -> bool:
        if self.start_len is None:
            self.start_len = self.input_ids.shape[1]
            return False
        last = int(output_ids[0, -1])
        if any(last == k for k in self.keyword_ids):
            return True
        text = self.tokenizer.batch_decode(output_ids[:, self.start_len:], skip_special_tokens=True)[0]
        return any(k in text for k in self.keywords)


def _training_only(*args, **kwargs):
    raise NotImplementedError("tokenizer/embedding resizing is a training-time helper; it is outside the "
                              "B200 inference hot path (use the reference implementation to prepare checkpoints)")


smart_resize_token_embedding = _training_only
smart_special_token_and_embedding_resize = _training_only
multi_modal_resize_token_embedding = _training_only
