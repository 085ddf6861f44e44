"""Host driver configuration (csrc/host/sim.cpp) without a GPU: Case.xml is parsed before any device call, so what the driver
accepts and what it refuses -- loudly, never silently -- can be checked on the CPU.  Everything refused here is a documented
limit of DESIGN.md section 7; everything accepted must get as far as the device (B2S_ERR_NODEVICE in this container, a running
simulation on the GPU box)."""
import pytest

import helpers as H
import dualsphysics_b200 as dsb

MOTION_OK = """        <begin mov="1" start="0" />
        <mvrect id="1" duration="0.1" next="2"><vel x="0.1" y="0" z="0" /></mvrect>
        <wait id="2" duration="0.1" next="3" />
        <mvrotsinu id="3" duration="1" next="1" anglesunits="radians"><freq v="2" /><ampl v="0.05" /><phase v="0" />
          <axisp1 x="0.1" y="0" z="0" /><axisp2 x="0.1" y="1" z="0" /></mvrotsinu>"""


def _load(tmp_path, options="", **spec):
    H.write_case(tmp_path, **spec)
    try:
        sim = dsb.Simulation(str(tmp_path / "Case"), options=options)
    except dsb.B2SError as e:
        return str(e)
    sim.close()
    return "loaded"


def _accepted(msg):
    # parsed and configured: either it runs (GPU present) or the only complaint is the missing device
    return msg == "loaded" or "no CUDA device" in msg or "b2s_create failed" in msg


REFUSED = [
    (dict(special_xml="<shifting><mode value=\"1\" /><shiftingzone /></shifting>"), "", "shifting zones"),
    (dict(params={"Shifting": 4}), "", "advanced shifting"),
    (dict(special_xml="<damping><dampingcylinder><overlimit value=\"1\" /></dampingcylinder></damping>"), "", "dampingcylinder"),
    (dict(special_xml="<gauges><mesh name=\"m\" /></gauges>"), "", "mesh"),
    (dict(piston=True, motion_xml="<begin mov=\"1\" start=\"0\" /><mvcir id=\"1\" duration=\"1\" />"), "", "mvcir"),
    (dict(piston=True, motion_xml="<begin mov=\"1\" start=\"0\" /><mvrectace id=\"1\" duration=\"1\"><ace x=\"1\" y=\"0\" z=\"0\" /></mvrectace>"), "", "velini"),
    (dict(floating_box=(0.15, 0.25, 0.2, 0.1, 0.15, 0.1), open_axes="y", params={"YPeriodicIncZ": 0}), "", "floating"),
    (dict(), "-shifting:fulladv", "shifting mode"),
    (dict(), "-dem", "not supported"),
    (dict(), "-b2svariant:2", "b2svariant"),
    (dict(params={"RigidAlgorithm": 2}), "", "RigidAlgorithm"),
    (dict(layers=4, normals=True, floating_box=(0.15, 0.25, 0.2, 0.1, 0.15, 0.1), params={"Boundary": 2, "SlipMode": 2}), "", "floating"),
]


@pytest.mark.parametrize("spec,options,needle", REFUSED, ids=[r[2].replace(" ", "_") for r in REFUSED])
def test_unsupported_configuration_is_refused_with_a_message(spec, options, needle, tmp_path, built_lib):
    msg = _load(tmp_path, options=options, dp=0.05, **spec)
    assert not _accepted(msg), f"accepted: {msg}"
    assert needle.lower() in msg.lower(), msg


ACCEPTED = [
    (dict(params={"Shifting": 2, "ShiftCoef": -2, "ShiftTFS": 1.5}), ""),
    (dict(), "-shifting:nobound -cellmode:half"),
    (dict(), "-b2svariant:5 -stable"),
    (dict(special_xml="<damping><dampingzone><limitmin x=\"1\" y=\"0\" z=\"0\" /><limitmax x=\"1.4\" y=\"0\" z=\"0\" /><overlimit value=\"0.1\" /></dampingzone>"
                      "<dampingbox><limitmin><pointini x=\"0.2\" y=\"0.2\" z=\"0.1\" /><pointend x=\"0.4\" y=\"0.4\" z=\"0.2\" /></limitmin>"
                      "<limitmax><pointini x=\"0.1\" y=\"0.1\" z=\"0\" /><pointend x=\"0.5\" y=\"0.5\" z=\"0.3\" /></limitmax><overlimit value=\"0.05\" />"
                      "<directions value=\"top,-bottom,left\" /></dampingbox></damping>"), ""),
    (dict(open_axes="y", params={"YPeriodicIncZ": 0}), ""),
    (dict(open_axes="xy", params={"XYPeriodic": 1}), ""),
    (dict(piston=True, motion_xml=MOTION_OK), ""),
    (dict(layers=4, normals=True, piston=True, params={"Boundary": 2, "SlipMode": 2}), ""),
    (dict(special_xml="<gauges><velocity name=\"v\"><point x=\"0.3\" y=\"0.3\" z=\"0.1\" /></velocity><force name=\"f\"><target mkbound=\"0\" /></force>"
                      "<maxz name=\"z\"><point0 x=\"0.3\" y=\"0.3\" z=\"0\" /><height value=\"0.5\" /><distlimit coefdp=\"3\" /></maxz></gauges>"), ""),
]


@pytest.mark.parametrize("spec,options", ACCEPTED, ids=[str(i) for i in range(len(ACCEPTED))])
def test_supported_configuration_reaches_the_device(spec, options, tmp_path, built_lib):
    msg = _load(tmp_path, options=options, dp=0.05, **spec)
    assert _accepted(msg), msg
