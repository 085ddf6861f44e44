"""The driver's XML reader (csrc/host/xml_lite.cpp) against Python's ElementTree: same number of elements and attributes on the
well-formed XML files the reference ships (the two example case definitions with their comments, headers and nested geometry
sections, DsphConfig.xml; the doc/xml_format templates are annotated text, not XML; the tree is present in the build container
only, skipped elsewhere) and on the case files this repository generates."""
import ctypes as C
import glob
import os
import xml.etree.ElementTree as ET

import pytest

import helpers as H
from dualsphysics_b200 import lib as L

REFXML = sorted(glob.glob("/root/reference/doc/**/*.xml", recursive=True) + glob.glob("/root/reference/examples/**/*.xml", recursive=True)
                + glob.glob("/root/reference/bin/linux/*.xml"))


def _stats(path):
    lib = L.load_library()
    ne, na = C.c_uint32(), C.c_uint32()
    r = lib.b2s_host_xml_stats(path.encode(), C.byref(ne), C.byref(na))
    assert r == 0, (path, lib.b2s_host_last_error().decode())
    return ne.value, na.value


def _et(path):
    root = ET.parse(path).getroot()
    nodes = list(root.iter())
    return len(nodes), sum(len(n.attrib) for n in nodes)


@pytest.mark.skipif(not REFXML, reason="the reference tree is only present in the build container")
def test_reader_agrees_with_elementtree_on_the_reference_files(built_lib):
    done = 0
    for f in REFXML:
        try:
            want = _et(f)
        except ET.ParseError:
            continue
        assert _stats(f) == want, f
        done += 1
    assert done >= 3


def test_reader_agrees_on_generated_cases_and_reports_errors(tmp_path, built_lib):
    H.write_case(tmp_path, dp=0.05, piston=True, floating_box=(0.15, 0.25, 0.2, 0.1, 0.15, 0.1))
    f = str(tmp_path / "Case.xml")
    assert _stats(f) == _et(f)
    bad = tmp_path / "bad.xml"
    bad.write_text("<case><a x='1'></b></case>")
    lib = L.load_library()
    assert lib.b2s_host_xml_stats(str(bad).encode(), None, None) != 0 and lib.b2s_host_last_error()
    assert lib.b2s_host_xml_stats(str(tmp_path / "missing.xml").encode(), None, None) != 0
