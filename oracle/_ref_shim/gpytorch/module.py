"""gpytorch.Module: nn.Module + named priors + named constraints + ``initialize`` (gpytorch/module.py semantics)."""
from collections import OrderedDict

import torch
from torch import nn


class Module(nn.Module):
    def __init__(self):
        super().__init__()
        self._added_loss_terms = OrderedDict()
        self._priors = OrderedDict()
        self._constraints = OrderedDict()

    def __call__(self, *inputs, **kwargs):
        outputs = self.forward(*inputs, **kwargs)
        if isinstance(outputs, list):
            return [o for o in outputs]
        return outputs

    # -- traversal ------------------------------------------------------------------------------
    def named_priors(self, memo=None, prefix=""):
        """(full name, owning module, prior, closure, setting closure): a module's own priors in registration order,
        then every child's (any nn.Module, e.g. the ModuleList of a ProductKernel) in ``named_children`` order."""
        return _extract_named_priors(self, memo=memo, prefix=prefix)

    def named_constraints(self, memo=None, prefix=""):
        return _extract_named_constraints(self, memo=memo, prefix=prefix)

    def named_hyperparameters(self):
        for name, param in self.named_parameters():
            yield name, param

    def hyperparameters(self):
        for _, param in self.named_hyperparameters():
            yield param

    # -- registration -----------------------------------------------------------------------------
    def register_parameter(self, name, parameter):
        if "_parameters" not in self.__dict__:
            raise AttributeError("Cannot assign parameter before Module.__init__() call")
        super().register_parameter(name, parameter)

    def register_prior(self, name, prior, param_or_closure, setting_closure=None):
        if isinstance(param_or_closure, str):
            if param_or_closure not in self._parameters and not hasattr(self, param_or_closure):
                raise AttributeError("Unknown parameter {name} for {module}".format(
                    name=param_or_closure, module=self.__class__.__name__))

            def closure(module, _name=param_or_closure):
                return getattr(module, _name)

            if setting_closure is not None:
                raise RuntimeError("Must specify a closure instead of a parameter name when providing setting_closure")

            def setting_closure(module, val, _name=param_or_closure):
                return module.initialize(**{_name: val})
        else:
            closure = param_or_closure
        self.add_module(name, prior)
        self._priors[name] = (prior, closure, setting_closure)

    def register_constraint(self, param_name, constraint, replace=True):
        if param_name not in self._parameters:
            raise RuntimeError("Attempting to register constraint for nonexistent parameter.")
        constraint_name = param_name + "_constraint"
        if constraint_name in self._constraints:
            current_constraint = self._constraints[constraint_name]
        else:
            current_constraint = None
        if isinstance(current_constraint, type(constraint)) and not replace:
            new_constraint = constraint.intersect(current_constraint)
        else:
            new_constraint = constraint
        self.add_module(constraint_name, new_constraint)
        self._constraints[constraint_name] = new_constraint
        if new_constraint.initial_value is not None:
            self.initialize(**{param_name: new_constraint.inverse_transform(new_constraint.initial_value)})

    def constraint_for_parameter_name(self, param_name):
        base_module = self
        base_name = param_name
        while "." in base_name:
            components = base_name.split(".")
            submodule_name = components[0]
            submodule = getattr(base_module, submodule_name)
            base_module = submodule
            base_name = ".".join(components[1:])
        try:
            constraint_name = base_name + "_constraint"
            return base_module._constraints.get(constraint_name)
        except AttributeError:
            return None

    # -- initialise (raw or constrained) values by (dotted) name -----------------------------------
    def initialize(self, **kwargs):
        for name, val in kwargs.items():
            if isinstance(val, int):
                val = float(val)
            if "." in name:
                module, name = self._get_module_and_name(name)
                if isinstance(module, nn.ModuleList):
                    idx, name = name.split(".", 1)
                    module[int(idx)].initialize(**{name: val})
                else:
                    module.initialize(**{name: val})
            elif not hasattr(self, name):
                raise AttributeError("Unknown parameter {p} for {c}".format(p=name, c=self.__class__.__name__))
            elif name not in self._parameters and name not in self._buffers:
                setattr(self, name, val)  # a property with a setter (constrained value)
            elif torch.is_tensor(val):
                constraint = self.constraint_for_parameter_name(name)
                if constraint is not None and constraint.enforced and not constraint.check_raw(val):
                    raise RuntimeError("Attempting to manually set a parameter value that is out of bounds of "
                                       "its current constraints, {}.".format(constraint))
                try:
                    self.__getattr__(name).data.copy_(val.expand_as(self.__getattr__(name)))
                except RuntimeError:
                    if not self.__getattr__(name).shape == val.shape:
                        raise
                    self.__getattr__(name).data.copy_(val.view_as(self.__getattr__(name)))
            elif isinstance(val, float):
                constraint = self.constraint_for_parameter_name(name)
                if constraint is not None and not constraint.check_raw(val):
                    raise RuntimeError("Attempting to manually set a parameter value that is out of bounds of "
                                       "its current constraints, {}.".format(constraint))
                self.__getattr__(name).data.fill_(val)
            else:
                raise AttributeError("Type {t} not valid for initializing parameter {p}".format(t=type(val), p=name))
            prior_name = "_".join([name, "prior"])
            if prior_name in self._priors:
                prior, closure, _ = self._priors[prior_name]
                try:
                    prior._validate_sample(closure(self))
                except ValueError as e:
                    raise ValueError("Invalid input value for prior {}. Error:\n{}".format(prior_name, e))
        return self

    def _get_module_and_name(self, parameter_name):
        module, name = parameter_name.split(".", 1)
        if module in self._modules:
            return self.__getattr__(module), name
        raise AttributeError("Invalid parameter name {}. {} has no module {}".format(
            parameter_name, type(self).__name__, module))

    def added_loss_terms(self):
        return iter(())


def _extract_named_priors(module, memo=None, prefix=""):
    if memo is None:
        memo = set()
    if hasattr(module, "_priors"):
        for name, (prior, closure, inv_closure) in module._priors.items():
            if prior is not None and prior not in memo:
                memo.add(prior)
                full_name = ("." if prefix else "").join([prefix, name])
                yield full_name, module, prior, closure, inv_closure
    for mname, module_ in module.named_children():
        submodule_prefix = prefix + ("." if prefix else "") + mname
        for item in _extract_named_priors(module_, memo=memo, prefix=submodule_prefix):
            yield item


def _extract_named_constraints(module, memo=None, prefix=""):
    if memo is None:
        memo = set()
    if hasattr(module, "_constraints"):
        for name, constraint in module._constraints.items():
            if constraint is not None and constraint not in memo:
                memo.add(constraint)
                full_name = ("." if prefix else "").join([prefix, name])
                yield full_name, constraint
    for mname, module_ in module.named_children():
        submodule_prefix = prefix + ("." if prefix else "") + mname
        for item in _extract_named_constraints(module_, memo=memo, prefix=submodule_prefix):
            yield item
