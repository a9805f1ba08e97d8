"""Model surgery helpers with the API of the reference's `utils/nethook.py`.

* `subsequence(sequential, first_layer=..., last_layer=..., after_layer=..., upto_layer=...,
  single_layer=..., share_weights=False)` slices nested `nn.Sequential`s by dotted names
  (reference: nethook.py:322-401).  With `share_weights=True` a child that is included whole
  is the ORIGINAL module object; every level that has to be entered becomes a plain
  `nn.Sequential` of the selected children — so the rewriter's context / target / rendering
  models share parameters by identity with the full generator.
* `InstrumentedModel` retains / edits layer outputs by installing an instance-level
  `forward` on the hooked module (reference: nethook.py:16-281).  The fused
  `StyledConvSeq.forward` of this package detects such instance-level forwards and drops to
  per-child execution, so hooks keep firing.
* `set_requires_grad(flag, *models_or_tensors)`.

No kernels here: this is the API surface the CUDA path plugs into.
"""
import copy
import inspect
import types
from collections import OrderedDict, defaultdict

import numpy
import torch


# ------------------------------------------------------------------------------------------
# subsequence
# ------------------------------------------------------------------------------------------
def subsequence(sequential, first_layer=None, last_layer=None, after_layer=None,
                upto_layer=None, single_layer=None, share_weights=False):
    """Sub-network from `first_layer` to `last_layer` inclusive, or strictly between
    `after_layer` and `upto_layer`.  Dotted names descend into nested Sequentials."""
    if single_layer is not None:
        assert first_layer is None and last_layer is None
        assert after_layer is None and upto_layer is None
        first_layer = last_layer = single_layer
    split = lambda s: None if s is None else s.split('.')
    return hierarchical_subsequence(sequential, first=split(first_layer), last=split(last_layer),
                                    after=split(after_layer), upto=split(upto_layer),
                                    share_weights=share_weights)


def hierarchical_subsequence(sequential, first, last, after, upto, share_weights=False, depth=0):
    """Recursive worker for `subsequence`; the four bounds are lists of name parts (or None
    once they no longer constrain this subtree)."""
    assert last is None or upto is None
    assert first is None or after is None
    if first is None and last is None and after is None and upto is None:
        return sequential if share_weights else copy.deepcopy(sequential)
    if not isinstance(sequential, torch.nn.Sequential):
        bound = first or last or after or upto
        raise AssertionError('.'.join(bound[:depth] or ['arg']) + ' not Sequential')

    def head(parts):
        """(name at this depth, True if the bound continues below this level)."""
        if parts is None:
            return None, False
        return parts[depth], len(parts) > depth + 1

    f_name, f_deep = head(first)
    l_name, l_deep = head(last)
    a_name, a_deep = head(after)
    u_name, u_deep = head(upto)
    pending = {'first': first is not None, 'last': last is not None,
               'after': after is not None, 'upto': upto is not None}

    taking = first is None and after is None
    chosen = OrderedDict()
    for name, child in sequential._modules.items():
        # bounds that open at this child
        if name == f_name:
            pending['first'] = False
            taking = True
        if name == a_name and a_deep:
            pending['after'] = False
            taking = True          # part of this child (after the nested bound) is wanted
        # an exclusive upper bound that names this child itself closes before it
        if name == u_name and not u_deep:
            pending['upto'] = False
            taking = False
        if taking:
            sub = hierarchical_subsequence(
                child,
                first=first if (f_deep and name == f_name) else None,
                last=last if (l_deep and name == l_name) else None,
                after=after if (a_deep and name == a_name) else None,
                upto=upto if (u_deep and name == u_name) else None,
                share_weights=share_weights, depth=depth + 1)
            if sub is not None:
                chosen[name] = sub
        # bounds that close after this child
        if name == l_name:
            pending['last'] = False
            taking = False
        if name == u_name and u_deep:
            pending['upto'] = False
            taking = False
        if name == a_name and not a_deep:
            pending['after'] = False
            taking = True
    for key, parts in (('first', first), ('last', last), ('after', after), ('upto', upto)):
        if pending[key]:
            raise ValueError('Layer %s not found' % '.'.join(parts))
    if not chosen and depth > 0:
        return None          # empty nested slices vanish; the outermost never returns None
    return torch.nn.Sequential(chosen)


def set_requires_grad(requires_grad, *models):
    for model in models:
        if isinstance(model, torch.nn.Module):
            for param in model.parameters():
                param.requires_grad = requires_grad
        elif isinstance(model, (torch.nn.Parameter, torch.Tensor)):
            model.requires_grad = requires_grad
        else:
            raise AssertionError('unknown type %r' % type(model))


# ------------------------------------------------------------------------------------------
# InstrumentedModel
# ------------------------------------------------------------------------------------------
def _name_and_aka(spec):
    if isinstance(spec, str):
        return spec, spec
    name, aka = spec
    return name, aka


class InstrumentedModel(torch.nn.Module):
    """Wraps a model so that named layers can be observed (`retain_layer`) or modified
    (`edit_layer`) on every forward pass.

        with InstrumentedModel(model) as inst:
            inst.retain_layer('layer4')
            inst.edit_layer('layer4', ablation=0.5, replacement=feats)
            inst(z)
            acts = inst.retained_layer('layer4')
    """

    def __init__(self, model):
        super().__init__()
        self.model = model
        self._retained = OrderedDict()
        self._detach_retained = {}
        self._editargs = defaultdict(dict)
        self._editrule = {}
        self._hooked_layer = {}
        self._old_forward = {}
        if isinstance(model, torch.nn.Sequential):
            self._hook_sequential()

    def __enter__(self):
        return self

    def __exit__(self, type, value, traceback):
        self.close()

    def forward(self, *inputs, **kwargs):
        return self.model(*inputs, **kwargs)

    def layer_names(self):
        return [name for name, _ in self.model.named_modules()]

    # -- retaining ---------------------------------------------------------------------------
    def retain_layer(self, layername, detach=True):
        self.retain_layers([layername], detach=detach)

    def retain_layers(self, layernames, detach=True):
        self.add_hooks(layernames)
        for spec in layernames:
            _, aka = _name_and_aka(spec)
            if aka not in self._retained:
                self._retained[aka] = None
                self._detach_retained[aka] = detach

    def stop_retaining_layers(self, layernames):
        self.add_hooks(layernames)
        for spec in layernames:
            _, aka = _name_and_aka(spec)
            if aka in self._retained:
                del self._retained[aka]
                del self._detach_retained[aka]

    def retained_features(self, clear=False):
        result = OrderedDict(self._retained)
        if clear:
            for k in result:
                self._retained[k] = None
        return result

    def retained_layer(self, aka=None, clear=False):
        if aka is None:
            aka = next(iter(self._retained))
        result = self._retained[aka]
        if clear:
            self._retained[aka] = None
        return result

    # -- editing -----------------------------------------------------------------------------
    def edit_layer(self, layername, rule=None, **kwargs):
        """Modify a layer's output on every run.  Default rule: `x*(1-a) + r*a` with
        keyword buffers `ablation=a`, `replacement=r`."""
        layername, aka = _name_and_aka(layername)
        if rule is None:
            rule = apply_ablation_replacement
        self.add_hooks([(layername, aka)])
        self._editargs[aka].update(kwargs)
        self._editrule[aka] = rule

    def remove_edits(self, layername=None):
        if layername is None:
            self._editargs.clear()
            self._editrule.clear()
            return
        _, aka = _name_and_aka(layername)
        self._editargs.pop(aka, None)
        self._editrule.pop(aka, None)

    # -- hook plumbing -----------------------------------------------------------------------
    def add_hooks(self, layernames):
        wanted = {}
        for spec in layernames:
            name, aka = _name_and_aka(spec)
            if self._hooked_layer.get(aka, None) != name:
                wanted[name] = aka
        if not wanted:
            return
        for name, layer in self.model.named_modules():
            if name in wanted:
                self._hook_layer(layer, name, wanted.pop(name))
        for name in wanted:
            raise ValueError('Layer %s not found in model' % name)

    def _hook_layer(self, layer, layername, aka):
        if aka in self._hooked_layer:
            raise ValueError('Layer %s already hooked' % aka)
        if layername in self._old_forward:
            raise ValueError('Layer %s already hooked' % layername)
        self._hooked_layer[aka] = layername
        self._old_forward[layername] = (layer, aka, layer.__dict__.get('forward', None))
        editor = self
        inner = layer.forward

        def hooked_forward(this, *inputs, **kwargs):
            return editor._postprocess_forward(inner(*inputs, **kwargs), aka)
        layer.forward = types.MethodType(hooked_forward, layer)

    def _unhook_layer(self, aka):
        if aka not in self._hooked_layer:
            return
        layername = self._hooked_layer[aka]
        if aka in self._retained:
            del self._retained[aka]
            del self._detach_retained[aka]
        self.remove_edits(aka)
        layer, check, old_forward = self._old_forward[layername]
        assert check == aka
        if old_forward is None:
            layer.__dict__.pop('forward', None)
        else:
            layer.forward = old_forward
        del self._old_forward[layername]
        del self._hooked_layer[aka]

    def _postprocess_forward(self, x, aka):
        if aka in self._retained:
            self._retained[aka] = x.detach() if self._detach_retained[aka] else x
        rule = self._editrule.get(aka, None)
        if rule is not None:
            x = invoke_with_optional_args(rule, x, self, name=aka, **(self._editargs[aka]))
        return x

    def _hook_sequential(self):
        """A Sequential root additionally accepts `layer=`, `first_layer=`, `last_layer=`
        keyword arguments to run only part of its children."""
        model = self.model
        self._hooked_layer['.'] = '.'
        self._old_forward['.'] = (model, '.', model.__dict__.get('forward', None))

        def partial_forward(this, x, layer=None, first_layer=None, last_layer=None):
            assert layer is None or (first_layer is None and last_layer is None)
            if layer is not None:
                first_layer = last_layer = layer
            first = None if first_layer is None else str(first_layer)
            last = None if last_layer is None else str(last_layer)
            running = first is None
            for name, child in this._modules.items():
                if name == first:
                    first, running = None, True
                if running:
                    x = child(x)
                if name == last:
                    last, running = None, False
            assert first is None, '%s not found' % first
            assert last is None, '%s not found' % last
            return x
        model.forward = types.MethodType(partial_forward, model)

    def close(self):
        for aka in list(self._hooked_layer.keys()):
            self._unhook_layer(aka)
        assert len(self._old_forward) == 0


def apply_ablation_replacement(x, imodel, **buffers):
    if buffers is not None:
        a = make_matching_tensor(buffers, 'ablation', x)
        if a is not None:
            x = x * (1 - a)
            v = make_matching_tensor(buffers, 'replacement', x)
            if v is not None:
                x += (v * a)
    return x


def make_matching_tensor(valuedict, name, data):
    """valuedict[name] as a tensor of data's dtype/device/rank (cached back into the dict)."""
    v = valuedict.get(name, None)
    if v is None:
        return None
    if not isinstance(v, torch.Tensor):
        v = torch.from_numpy(numpy.array(v))
        valuedict[name] = v
    if v.device != data.device or v.dtype != data.dtype:
        assert not v.requires_grad, '%s wrong device or type' % name
        v = v.to(device=data.device, dtype=data.dtype)
        valuedict[name] = v
    if len(v.shape) < len(data.shape):
        assert not v.requires_grad, '%s wrong dimensions' % name
        v = v.view((1,) + tuple(v.shape) + (1,) * (len(data.shape) - len(v.shape) - 1))
        valuedict[name] = v
    return v


def invoke_with_optional_args(fn, *args, **kwargs):
    """Call fn with as many of the positional / keyword arguments as its signature takes."""
    spec = inspect.getfullargspec(fn)
    taken = 0
    if spec.varkw is None:
        taken = len([k for k in kwargs if k in spec.args])
        kwargs = {k: v for k, v in kwargs.items()
                  if k in spec.args or (spec.kwonlyargs and k in spec.kwonlyargs)}
    if spec.varargs is None:
        args = args[:len(spec.args) - taken]
    return fn(*args, **kwargs)
