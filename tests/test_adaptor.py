"""The C++ drop-in adaptor (include/wso_tessendorf_adaptor.hpp) and the headless caller (examples/frame_loop.cpp):
compile checks on CPU, end-to-end run against the Python path on the GPU."""
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

EX = os.path.join(ROOT, "examples")
GXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def _build_frame_loop():
    from watersurfacerendering_b200 import build as B
    B.build()
    subprocess.check_call(["make", "-C", EX, "frame_loop"], stdout=subprocess.DEVNULL)
    return os.path.join(EX, "frame_loop")


def test_adaptor_and_caller_compile():
    exe = _build_frame_loop()
    assert os.path.exists(exe)


def test_adaptor_exposes_the_reference_interface(tmp_path):
    """Every public member the reference's caller uses (WaterSurfaceMesh.cpp:127-179, 705-741, 804-902)."""
    src = tmp_path / "iface.cpp"
    src.write_text(r'''
#include <cstring>
#include <vector>
#include "wso_tessendorf_adaptor.hpp"
static_assert(sizeof(WSTessendorf::Displacement) == 16 && sizeof(WSTessendorf::Normal) == 16, "RGBA32F texels");
float use(WSTessendorf& m, std::vector<unsigned char>& staging) {
    m.SetTileSize(WSTessendorf::s_kDefaultTileSize);
    m.SetTileLength(WSTessendorf::s_kDefaultTileLength);
    m.SetWindDirection(WSTessendorf::s_kDefaultWindDir);
    m.SetWindSpeed(WSTessendorf::s_kDefaultWindSpeed);
    m.SetAnimationPeriod(WSTessendorf::s_kDefaultAnimPeriod);
    m.SetPhillipsConst(WSTessendorf::s_kDefaultPhillipsConst);
    m.SetDamping(WSTessendorf::s_kDefaultPhillipsDamping);
    m.SetLambda(-1.0f);
    m.Prepare();
    float a = m.ComputeWaves(0.5f);
    const size_t dbytes = sizeof(WSTessendorf::Displacement) * m.GetDisplacementCount();
    const size_t nbytes = sizeof(WSTessendorf::Normal) * m.GetNormalCount();
    staging.resize(dbytes + nbytes);
    std::memcpy(staging.data(), m.GetDisplacements().data(), dbytes);
    std::memcpy(staging.data() + dbytes, m.GetNormals().data(), nbytes);
    auto w = m.GetWindDir();
    return a + m.GetTileSize() + m.GetTileLength() + w.x + w.y + m.GetWindSpeed() + m.GetAnimationPeriod() +
           m.GetPhillipsConst() + m.GetDamping() + m.GetDisplacementLambda() + m.GetMinHeight() + m.GetMaxHeight();
}
''')
    subprocess.check_call([GXX, "-std=c++17", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), str(src)])
    glm = "/root/reference/libs/glm"
    if os.path.isdir(glm):   # with glm in scope the adaptor uses glm::vec2 / glm::vec4 like the reference
        src2 = tmp_path / "iface_glm.cpp"
        src2.write_text('#include <glm/glm.hpp>\n' + src.read_text() +
                        '\nstatic_assert(std::is_same<WSTessendorf::Displacement, glm::vec4>::value, "glm types");\n')
        subprocess.check_call([GXX, "-std=c++17", "-fsyntax-only", "-I", glm, "-I", os.path.join(ROOT, "include"),
                               str(src2)])


REF_CALLER = "/root/reference/src/scene/WaterSurfaceMesh.cpp"


@pytest.mark.skipif(not os.path.exists(REF_CALLER), reason="the reference tree exists in the build container only")
def test_reference_callers_own_expressions_compile_against_the_adaptor(tmp_path):
    """Not a hand-written list: every `m_ModelTess->...` expression, the constructor call and every `WSTessendorf::` name
    is cut out of the reference's caller as it lies under /root/reference (WaterSurfaceMesh.cpp) and compiled against the
    adaptor, with the caller's own variable types (WaterSurfaceMesh.h / .cpp: float times and sliders, glm::vec2 wind,
    uint32_t sizes) and glm in scope as in the reference build.  The Vulkan / ImGui code around them is not needed for that."""
    import re
    src = open(REF_CALLER).read()
    calls = [m.group(0) for m in re.finditer(r"m_ModelTess->\w+\s*\([^;]*?\)(?:\.data\(\))?(?=\s*[;,)*])", src)]
    ctor = re.findall(r"m_ModelTess\.reset\(\s*new (WSTessendorf\([^;]*?\))\s*\)\s*;", src)
    names = sorted(set(re.findall(r"WSTessendorf::(\w+)", src)) | set(re.findall(r"m_ModelTess->(s_k\w+)", src)))
    members = {c.split("->")[1].split("(")[0].strip() for c in calls}
    assert {"Prepare", "ComputeWaves", "GetDisplacements", "GetNormals", "SetTileSize", "SetLambda"} <= members
    assert len(calls) >= 30 and len(ctor) == 1, (len(calls), ctor)
    body = []
    for i, c in enumerate(calls):
        c = " ".join(c.split())
        body.append(f"    {{ auto&& r{i} = ({c}, 0); (void)r{i}; }}" if re.search(r"->(Set\w+|Prepare)\(", c)
                    else f"    {{ auto&& r{i} = {c}; (void)r{i}; }}")
    types = "\n".join(f"    (void)sizeof(WSTessendorf::{n});" for n in names)
    tu = tmp_path / "ref_calls.cpp"
    tu.write_text(f"""
#include <cstdint>
#include <memory>
#include <glm/glm.hpp>
#include "wso_tessendorf_adaptor.hpp"
void reference_call_sites() {{
    // the caller's variables (WaterSurfaceMesh.h: m_TimeCtr; WaterSurfaceMesh.cpp:482-483, 804-890)
    float m_TimeCtr = 0.f, tileLen = 1000.f, windSpeed = 30.f, animPeriod = 200.f, phillipsA = 3.f, damping = 0.1f, lambda = -1.f;
    glm::vec2 windDir(1.f, 1.f);
    const uint32_t kNewSize = 256;
    const uint32_t s_kWSResolutions[7] = {{16, 32, 64, 128, 256, 512, 1024}};
    int tileRes = 5;
    const auto kSampleCount = WSTessendorf::s_kDefaultTileSize;
    const auto kWaveLength = WSTessendorf::s_kDefaultTileLength;
    std::unique_ptr<WSTessendorf> m_ModelTess;
    m_ModelTess.reset(new {ctor[0]});
{chr(10).join(body)}
{types}
    const glm::vec4* d = m_ModelTess->GetDisplacements().data();   // what vkp::Buffer::CopyToMapped receives (cpp:728, 741)
    const glm::vec4* n = m_ModelTess->GetNormals().data();
    (void)d; (void)n;
}}
""")
    subprocess.check_call([GXX, "-std=c++17", "-fsyntax-only", "-Wall", "-I", "/root/reference/libs/glm", "-I",
                           os.path.join(ROOT, "include"), str(tu)])


@pytest.mark.gpu
def test_frame_loop_matches_python_path():
    exe = _build_frame_loop()
    n, frames, seed = 256, 12, 4321
    out = subprocess.run([exe, str(n), str(frames), str(seed)], capture_output=True, text=True, check=True).stdout
    r = json.loads(out)
    import watersurfacerendering_b200 as W
    with W.WSTessendorf(n, 1000.0 * n / 512) as ws:
        ws.Prepare(seed=seed)
        t = np.float32(0.0)
        for _ in range(frames):
            t = np.float32(t + np.float32(np.float32(1.0 / 60.0) * np.float32(3.0)))
        a = ws.ComputeWaves(float(t))
        chk = float(ws.GetDisplacements().astype(np.float64).sum() + ws.GetNormals().astype(np.float64).sum())
        assert abs(r["t_last"] - float(t)) < 1e-6
        assert r["amplitude_last"] == pytest.approx(float(a), rel=1e-7)
        assert r["min_height"] == pytest.approx(float(ws.GetMinHeight()), rel=1e-7)
        assert r["checksum"] == pytest.approx(chk, rel=1e-9, abs=1e-3)
        assert r["tile_frames_per_s"] > 0
