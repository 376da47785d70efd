"""Restatement of the third-party package the reference's MMBT path depends on.
TEST INFRASTRUCTURE ONLY (never imported by the product).

The reference imports ``pytorch_pretrained_bert`` (PyPI ``pytorch-pretrained-bert``, HuggingFace
2019; **absent from the reference's requirements.txt, version unpinned, not installed here, no
network**) at ``src/mmbt.py:13`` (``BertModel``), ``train.py:16`` (``BertAdam``) and
``src/dataset.py:22`` (``BertTokenizer``).  This module restates that package's published
definitions -- release 0.6.2, the last one -- of exactly the classes those call sites touch
(``modeling.BertModel`` with ``embeddings`` / ``encoder`` / ``pooler``; ``optimization.BertAdam``)
so that ``tests/golden/make_golden.py`` can run the UNMODIFIED reference ``src/mmbt.py`` on top
of it and freeze golden vectors.  Parity for the MMBT path is therefore anchored on the
reference's own call sites (embedding assembly, mask construction, ``forward_control`` index
sampling, classifier, loss) while the BERT arithmetic is pinned to this restatement, not to
the absent package.  The restatement itself is held, in fp64 and to 1e-12, to
``transformers.BertModel`` (5.5.0, installed here: the maintained successor of the absent package,
same authors, same state-dict keys -- strict load) by
``tests/test_oracle_golden.py::test_restated_bert_matches_transformers_bertmodel``; ``BertAdam`` has
no surviving counterpart and stays pinned to its published definition only.

Facts restated (pytorch_pretrained_bert 0.6.2, modeling.py / optimization.py):
* ``gelu(x) = x * 0.5 * (1 + erf(x / sqrt(2)))``;
* ``BertLayerNorm``: TF style, ``(x - u) / sqrt(var + 1e-12)`` with the biased variance;
* ``BertEmbeddings``: word (``padding_idx=0``) + position (``arange(seq_len)``) + token type,
  LayerNorm, dropout;
* ``BertSelfAttention``: separate query / key / value Linears, scores / sqrt(head_dim) + additive
  mask, softmax, dropout, context; ``BertSelfOutput`` / ``BertOutput``: dense, dropout,
  ``LayerNorm(hidden + input)`` (post-LN); ``BertIntermediate``: dense + gelu;
* ``BertEncoder.forward(hidden, mask, output_all_encoded_layers=True)`` returns a list;
* ``BertPooler``: ``tanh(dense(hidden[:, 0]))``;
* ``init_bert_weights``: N(0, initializer_range) for Linear / Embedding weights, LayerNorm (1, 0),
  Linear biases 0;
* ``BertAdam.step``: per-parameter ``clip_grad_norm_(p, max_grad_norm)``, moments without bias
  correction, ``update = m / (sqrt(v) + e) + weight_decay * p``,
  ``lr_scheduled = lr * warmup_linear(step / t_total, warmup)``.
"""
import copy
import math

import torch
from torch import nn
from torch.nn.utils import clip_grad_norm_


def gelu(x):
    return x * 0.5 * (1.0 + torch.erf(x / math.sqrt(2.0)))


class BertConfig:
    def __init__(self, vocab_size=30522, hidden_size=768, num_hidden_layers=12,
                 num_attention_heads=12, intermediate_size=3072, hidden_dropout_prob=0.1,
                 attention_probs_dropout_prob=0.1, max_position_embeddings=512, type_vocab_size=2,
                 initializer_range=0.02):
        self.vocab_size = vocab_size
        self.hidden_size = hidden_size
        self.num_hidden_layers = num_hidden_layers
        self.num_attention_heads = num_attention_heads
        self.intermediate_size = intermediate_size
        self.hidden_dropout_prob = hidden_dropout_prob
        self.attention_probs_dropout_prob = attention_probs_dropout_prob
        self.max_position_embeddings = max_position_embeddings
        self.type_vocab_size = type_vocab_size
        self.initializer_range = initializer_range


class BertLayerNorm(nn.Module):
    def __init__(self, hidden_size, eps=1e-12):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(hidden_size))
        self.bias = nn.Parameter(torch.zeros(hidden_size))
        self.variance_epsilon = eps

    def forward(self, x):
        u = x.mean(-1, keepdim=True)
        s = (x - u).pow(2).mean(-1, keepdim=True)
        x = (x - u) / torch.sqrt(s + self.variance_epsilon)
        return self.weight * x + self.bias


class BertEmbeddings(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.word_embeddings = nn.Embedding(config.vocab_size, config.hidden_size, padding_idx=0)
        self.position_embeddings = nn.Embedding(config.max_position_embeddings, config.hidden_size)
        self.token_type_embeddings = nn.Embedding(config.type_vocab_size, config.hidden_size)
        self.LayerNorm = BertLayerNorm(config.hidden_size, eps=1e-12)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)

    def forward(self, input_ids, token_type_ids=None):
        seq_length = input_ids.size(1)
        position_ids = torch.arange(seq_length, dtype=torch.long, device=input_ids.device)
        position_ids = position_ids.unsqueeze(0).expand_as(input_ids)
        if token_type_ids is None:
            token_type_ids = torch.zeros_like(input_ids)
        embeddings = (self.word_embeddings(input_ids) + self.position_embeddings(position_ids)
                      + self.token_type_embeddings(token_type_ids))
        return self.dropout(self.LayerNorm(embeddings))


class BertSelfAttention(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.num_attention_heads = config.num_attention_heads
        self.attention_head_size = config.hidden_size // config.num_attention_heads
        self.all_head_size = self.num_attention_heads * self.attention_head_size
        self.query = nn.Linear(config.hidden_size, self.all_head_size)
        self.key = nn.Linear(config.hidden_size, self.all_head_size)
        self.value = nn.Linear(config.hidden_size, self.all_head_size)
        self.dropout = nn.Dropout(config.attention_probs_dropout_prob)

    def transpose_for_scores(self, x):
        x = x.view(*x.size()[:-1], self.num_attention_heads, self.attention_head_size)
        return x.permute(0, 2, 1, 3)

    def forward(self, hidden_states, attention_mask):
        q = self.transpose_for_scores(self.query(hidden_states))
        k = self.transpose_for_scores(self.key(hidden_states))
        v = self.transpose_for_scores(self.value(hidden_states))
        scores = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(self.attention_head_size)
        scores = scores + attention_mask
        probs = self.dropout(nn.Softmax(dim=-1)(scores))
        ctx = torch.matmul(probs, v).permute(0, 2, 1, 3).contiguous()
        return ctx.view(*ctx.size()[:-2], self.all_head_size)


class BertSelfOutput(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)
        self.LayerNorm = BertLayerNorm(config.hidden_size, eps=1e-12)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)

    def forward(self, hidden_states, input_tensor):
        return self.LayerNorm(self.dropout(self.dense(hidden_states)) + input_tensor)


class BertAttention(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.self = BertSelfAttention(config)
        self.output = BertSelfOutput(config)

    def forward(self, input_tensor, attention_mask):
        return self.output(self.self(input_tensor, attention_mask), input_tensor)


class BertIntermediate(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.intermediate_size)

    def forward(self, hidden_states):
        return gelu(self.dense(hidden_states))


class BertOutput(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.intermediate_size, config.hidden_size)
        self.LayerNorm = BertLayerNorm(config.hidden_size, eps=1e-12)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)

    def forward(self, hidden_states, input_tensor):
        return self.LayerNorm(self.dropout(self.dense(hidden_states)) + input_tensor)


class BertLayer(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.attention = BertAttention(config)
        self.intermediate = BertIntermediate(config)
        self.output = BertOutput(config)

    def forward(self, hidden_states, attention_mask):
        attention_output = self.attention(hidden_states, attention_mask)
        return self.output(self.intermediate(attention_output), attention_output)


class BertEncoder(nn.Module):
    def __init__(self, config):
        super().__init__()
        layer = BertLayer(config)
        self.layer = nn.ModuleList([copy.deepcopy(layer) for _ in range(config.num_hidden_layers)])

    def forward(self, hidden_states, attention_mask, output_all_encoded_layers=True):
        all_encoder_layers = []
        for layer_module in self.layer:
            hidden_states = layer_module(hidden_states, attention_mask)
            if output_all_encoded_layers:
                all_encoder_layers.append(hidden_states)
        if not output_all_encoded_layers:
            all_encoder_layers.append(hidden_states)
        return all_encoder_layers


class BertPooler(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)
        self.activation = nn.Tanh()

    def forward(self, hidden_states):
        return self.activation(self.dense(hidden_states[:, 0]))


class BertModel(nn.Module):
    #: ``from_pretrained(name)`` looks the configuration up here (there are no weight files
    #: offline: the golden script registers a small seeded configuration under a made-up name)
    CONFIGS = {"bert-base-uncased": BertConfig()}

    def __init__(self, config):
        super().__init__()
        self.config = config
        self.embeddings = BertEmbeddings(config)
        self.encoder = BertEncoder(config)
        self.pooler = BertPooler(config)
        self.apply(self.init_bert_weights)

    def init_bert_weights(self, module):
        if isinstance(module, (nn.Linear, nn.Embedding)):
            module.weight.data.normal_(mean=0.0, std=self.config.initializer_range)
        elif isinstance(module, BertLayerNorm):
            module.bias.data.zero_()
            module.weight.data.fill_(1.0)
        if isinstance(module, nn.Linear) and module.bias is not None:
            module.bias.data.zero_()

    @classmethod
    def from_pretrained(cls, name, *args, **kwargs):
        return cls(cls.CONFIGS[name])


def warmup_linear(x, warmup=0.002):
    if x < warmup:
        return x / warmup
    return 1.0 - x


class BertAdam(torch.optim.Optimizer):
    def __init__(self, params, lr, warmup=-1, t_total=-1, schedule="warmup_linear", b1=0.9, b2=0.999,
                 e=1e-6, weight_decay=0.01, max_grad_norm=1.0):
        defaults = dict(lr=lr, schedule=schedule, warmup=warmup, t_total=t_total, b1=b1, b2=b2, e=e,
                        weight_decay=weight_decay, max_grad_norm=max_grad_norm)
        super().__init__(params, defaults)

    def step(self, closure=None):
        loss = None if closure is None else closure()
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                grad = p.grad.data
                state = self.state[p]
                if len(state) == 0:
                    state["step"] = 0
                    state["next_m"] = torch.zeros_like(p.data)
                    state["next_v"] = torch.zeros_like(p.data)
                next_m, next_v = state["next_m"], state["next_v"]
                beta1, beta2 = group["b1"], group["b2"]
                if group["max_grad_norm"] > 0:
                    clip_grad_norm_(p, group["max_grad_norm"])
                next_m.mul_(beta1).add_(grad, alpha=1 - beta1)
                next_v.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
                update = next_m / (next_v.sqrt() + group["e"])
                if group["weight_decay"] > 0.0:
                    update += group["weight_decay"] * p.data
                if group["t_total"] != -1:
                    lr_scheduled = group["lr"] * warmup_linear(state["step"] / group["t_total"],
                                                               group["warmup"])
                else:
                    lr_scheduled = group["lr"]
                p.data.add_(-lr_scheduled * update)
                state["step"] += 1
        return loss
