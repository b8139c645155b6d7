"""The C-ABI library loads and exports every symbol include/stackrl_b200.h
declares (CPU only; no compute call is made)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'stackrl_b200.h')
LIB = os.path.join(ROOT, 'stackrl_b200', 'libstackrl_b200.so')


def declared_symbols():
  text = open(HEADER).read()
  # entry points still being built sit in an `#ifdef SRL_NEXT` block
  text = re.sub(r'#ifdef SRL_NEXT.*?#endif\s*/\* SRL_NEXT \*/', '', text, flags=re.S)
  text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
  return sorted(set(re.findall(r'SRL_API\s+[\w\s\*]+?\b(srl_\w+)\s*\(', text)))


@pytest.fixture(scope='module')
def lib():
  if not os.path.exists(LIB):
    from stackrl_b200 import build
    build.build()
  return ctypes.CDLL(LIB)


def test_header_declares_the_hot_path():
  names = declared_symbols()
  for required in ('srl_maxplus_f32', 'srl_version', 'srl_last_error'):
    assert required in names
  assert len(names) >= 4


def test_every_declared_symbol_is_exported(lib):
  missing = [n for n in declared_symbols() if not hasattr(lib, n)]
  assert not missing, 'declared in the header but not exported: {}'.format(missing)


def test_version_and_error_string(lib):
  assert lib.srl_version() >= 100
  lib.srl_last_error.restype = ctypes.c_char_p
  assert isinstance(lib.srl_last_error(), bytes)


def test_python_binding_covers_every_symbol(lib):
  from stackrl_b200 import capi
  assert set(declared_symbols()) <= set(capi._SIGNATURES)
