"""Recipe compiler: a dspeed DSP configuration (JSON / YAML / dict, the reference's schema) -> a built
:class:`~dspeed_b200.processing_chain.ProcessingChain`.

Three phases, each with its own data:

``parse``        every entry of ``processors`` becomes a :class:`Step` -- output names, callable (module + function) or
                 bare expression, argument texts with ``db.*`` tokens already replaced by database values / defaults,
                 and the names it reads;
``schedule``     a depth-first walk from the REQUESTED outputs orders the steps that are needed (nothing else is ever
                 instantiated), finds the table columns to read, and rejects circular references;
``instantiate``  in that order: variables with declared units, factory processors (``init_args``), constant folding (a
                 step whose inputs are all constants runs once now and its outputs become constants -- cusp / zac / t0
                 / gaussian kernels), everything else becomes a launch descriptor appended to the chain.

The schema and its semantics are the reference's (src/dspeed/processing_chain.py:2363-2872, docs/source/manuals):
an existing configuration compiles to the same plan (tests/test_chain_plan.py compares against the reference's own
config files).  The organisation -- typed steps, separate phases -- is this repository's.
"""
from __future__ import annotations

import ast
import json
import logging
import re
from collections.abc import MutableMapping
from copy import deepcopy
from dataclasses import dataclass, field

from .errors import ProcessingChainError

log = logging.getLogger("dspeed")

_DB_TOKEN = re.compile(r"(?![^\w_.])db\.[\w_.]+")
_NAME_SEP = re.compile(",| ")


def output_names(key: str) -> list[str]:
    """``"a, b"`` -> ``["a", "b"]``"""
    return [k for k in _NAME_SEP.split(key) if k != ""]


@dataclass
class Step:
    key: str                      # the entry's key as written ("tp_min, tp_max, wf_min, wf_max")
    outputs: list                 # its output variable names
    module: str | None            # None: the step is a bare expression (an alias, a constant, arithmetic)
    function: str
    args: list
    entry: dict                   # the original entry: unit, defaults, kwargs, init_args, lh5_attrs, description ...
    needs: list = field(default_factory=list)


class Database:
    """``db.a.b`` look-ups: the database dictionary first, the entry's ``defaults`` second"""

    def __init__(self, values):
        self.values = values

    def lookup(self, token: str, entry: dict):
        node = self.values
        try:
            for part in token[3:].split("."):
                node = node[part]
            return node
        except (KeyError, TypeError):
            pass
        try:
            return entry["defaults"][token]
        except (KeyError, TypeError):
            raise ProcessingChainError(f"did not find {token} in database, and could not find default value.")

    def substitute(self, text, entry: dict):
        """a text that IS a token becomes the value itself (any type); tokens inside a longer expression are spliced in"""
        for token in _DB_TOKEN.findall(text):
            value = self.lookup(token, entry)
            text = value if text == token else text.replace(token, str(value))
            if not isinstance(text, str):
                break
        return text


def _callable_of(text: str, entry: dict, key: str, expr_modules, expr_functions):
    """(module, function, args) named by an entry's ``function`` text.  ``module`` None: the text is an expression."""
    tree = ast.parse(text, mode="eval").body
    twice = f"Module specified twice for parameter {key}"
    no_args = f"Cannot specify arguments if function is expr for parameter {key}"

    def src(node):
        return text[node.col_offset: node.end_col_offset]

    def call_args(call):
        return [src(a) for a in call.args + call.keywords]

    if isinstance(tree, ast.Name):                       # "trap_norm" + separate module / args
        return entry.get("module", _MISSING), text, entry.get("args", _MISSING)
    if isinstance(tree, ast.Attribute):                  # "dspeed.processors.trap_norm" or an expression like np.pi
        module = src(tree.value)
        if module in expr_modules and "args" not in entry:
            return None, text, [text]
        if "module" in entry:
            raise ProcessingChainError(twice)
        return module, tree.attr, entry.get("args", _MISSING)
    if isinstance(tree, ast.Call):                       # the whole call in one string
        if "args" in entry:
            raise ProcessingChainError(no_args)
        if isinstance(tree.func, ast.Name):
            if tree.func.id in expr_functions and "module" not in entry:
                return None, text, [text]                # round(x, 4), where(...), len(...): grammar, not a processor
            return entry.get("module", _MISSING), tree.func.id, call_args(tree)
        if isinstance(tree.func, ast.Attribute):
            if "module" in entry:
                raise ProcessingChainError(twice)
            return src(tree.func.value), tree.func.attr, call_args(tree)
        return entry.get("module", _MISSING), text, entry.get("args", _MISSING)
    if "args" in entry:                                  # arithmetic, comparison, slicing, a if b else c ...
        raise ProcessingChainError(no_args)
    if "module" in entry:
        raise ProcessingChainError(twice)
    return None, text, [text]


_MISSING = object()


def parse(config, db_dict, names_read, expr_modules, expr_functions):
    """-> (steps by output name, requested outputs or None).  `names_read(text)` lists the variable names an argument
    expression reads (the chain's expression grammar decides what is a name)."""
    if isinstance(config, str):
        from yaml import safe_load

        with open(config) as f:
            config = safe_load(f)
    elif config is None:
        config = {}
    elif isinstance(config, MutableMapping):
        config = deepcopy(config)
    else:
        raise ValueError("processors must be a dict, json/yaml file, or None")
    requested = config.get("outputs") if "outputs" in config else None
    entries = dict(config["processors"]) if "processors" in config else dict(config)
    entries.pop("outputs", None) if "processors" not in config else None
    db = Database(db_dict)
    steps = {}
    for key, entry in entries.items():
        if isinstance(entry, str):
            entry = {"function": entry}
        if "function" not in entry:
            raise ProcessingChainError(f"no function given for parameter {key}")
        module, function, args = _callable_of(entry["function"], entry, key, expr_modules, expr_functions)
        if module is _MISSING:
            raise ProcessingChainError(f"Could not find module for parameter {key}")
        if args is _MISSING:
            raise ProcessingChainError(f"Could not find args for parameter {key}")
        args = [db.substitute(a, entry) if isinstance(a, str) else a for a in args]
        outs = output_names(key)
        if "prereqs" in entry:
            needs = list(entry["prereqs"])
        else:
            needs = []
            for a in args:
                if isinstance(a, str):
                    for name in names_read(a):
                        if name not in needs and name not in outs:
                            needs.append(name)
        # the entry as later phases (and error reports) see it: normalised in place like the reference does
        entry = dict(entry, function=function, module=module, args=args, prereqs=needs)
        step = Step(key, outs, module, function, args, entry, needs)
        log.debug(f"prereqs for {key} are {needs}")
        for name in outs if len(outs) > 1 else [key]:
            steps[name] = step
        steps[key] = step
    return steps, requested, db


def schedule(steps, outputs):
    """-> (steps to instantiate in dependency order, table columns they read, requested names that are plain copies of
    table columns, requested names that steps produce)"""
    ordered, columns, copies, produced = [], [], [], []
    done, open_ = set(), []

    def visit(name):
        step = steps.get(name)
        if step is None:
            if name not in columns:
                columns.append(name)
            return
        if id(step) in done:
            return
        if id(step) in open_:
            raise ProcessingChainError(f"Circular references detected for parameter '{name}'")
        open_.append(id(step))
        for need in step.needs:
            visit(need)
        open_.remove(id(step))
        done.add(id(step))
        ordered.append(step)

    for name in outputs:
        if name in steps:
            visit(name)
            produced.append(name)
        else:
            copies.append(name)
    return ordered, columns, copies, produced


def describe(step: Step) -> str:
    return json.dumps(step.entry, indent=2, default=str)
