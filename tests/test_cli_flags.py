"""The two command lines keep the reference's flags and defaults (build container only: the parsers of
scripts/neoloop-caller and scripts/assemble-complexSVs are loaded from the reference tree and compared)."""
import argparse
import importlib.machinery
import importlib.util
import os
import sys

import pytest

from oracle.run_reference import REFERENCE_ROOT, reference_available

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not reference_available(), reason="reference tree only exists in the build container")]


def _reference_parser(script, monkeypatch):
    """Run the reference script's getargs() with argparse instrumented to hand back the parser."""
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    path = os.path.join(REFERENCE_ROOT, "scripts", script)
    loader = importlib.machinery.SourceFileLoader("_ref_cli_" + script.replace("-", "_"), path)
    mod = importlib.util.module_from_spec(importlib.util.spec_from_loader(loader.name, loader))
    loader.exec_module(mod)
    seen = []
    real = argparse.ArgumentParser.parse_args

    def spy(self, *a, **k):
        seen.append(self)
        return real(self, *a, **k)
    monkeypatch.setattr(argparse.ArgumentParser, "parse_args", spy)
    monkeypatch.setattr(sys, "argv", [script, "-O", "x"])
    mod.getargs()
    return seen[0]


def _table(parser):
    out = {}
    for a in parser._actions:
        if a.option_strings and a.dest not in ("help", "version"):
            out[tuple(sorted(a.option_strings))] = (a.dest, a.default, a.nargs, tuple(a.choices) if a.choices else None,
                                                    getattr(a.type, "__name__", None), type(a).__name__)
    return out


def _mine(mod, monkeypatch):
    seen = []
    real = argparse.ArgumentParser.parse_args

    def spy(self, *a, **k):
        seen.append(self)
        return real(self, *a, **k)
    monkeypatch.setattr(argparse.ArgumentParser, "parse_args", spy)
    mod.getargs(["-O", "x"])
    return seen[0]


def test_neoloop_caller_flags(monkeypatch):
    from neoloopfinder_b200 import cli
    ref = _table(_reference_parser("neoloop-caller", monkeypatch))
    mine = _table(_mine(cli, monkeypatch))
    for key, spec in ref.items():
        assert mine.get(key) == spec, key
    assert {k for k in mine if k not in ref} == {("--model",), ("--gpus",), ("--genome-wide",), ("--chroms",)}   # documented additions


def test_assemble_complexsvs_flags(monkeypatch):
    from neoloopfinder_b200 import cli_assemble
    ref = _table(_reference_parser("assemble-complexSVs", monkeypatch))
    mine = _table(_mine(cli_assemble, monkeypatch))
    for key, spec in ref.items():
        assert mine.get(key) == spec, key
    assert {k for k in mine if k not in ref} == {("--protocol",)}                     # the flag the reference reads but never defines
