from rl4co.utils.ops import gather_by_index


def get_log_likelihood(logprobs, actions=None, mask=None, return_sum=True):
    if logprobs.dim() == 3:
        logprobs = gather_by_index(logprobs, actions, dim=-1)
    if mask is not None:
        logprobs[~mask] = 0
    assert (logprobs > -1000).data.all(), "Logprobs should not be -inf, check sampling procedure!"
    return logprobs.sum(1) if return_sum else logprobs
