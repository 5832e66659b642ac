"""The C++ host (minimc_b200/host) without a GPU: deck parsing, the object model's construction semantics and
error messages (the reference's test_Nuclide.cpp / test_World.cpp / test_Material.cpp / test_CSGSurface.cpp /
test_XMLDocument.cpp cases), flattening against the oracle's independent restatement and the reference's own
dump, and the .out text layout."""
import numpy as np
import pytest

import util
from minimc_b200 import capi, decks


def _arrays(world_json):
    out = {}
    for k, v in world_json.items():
        out[k] = [float.fromhex(x) if isinstance(x, str) else x for x in v] if isinstance(v, list) else v
    return out


@pytest.mark.parametrize("name", list(decks.DECKS))
def test_flatten_matches_oracle_restatement(name):
    text = util.deck_text(name, "surface")
    mine = _arrays(capi.Driver(text=text).world_json())
    ref = util.flat_from_xml(text)["world"]
    for key, value in ref.items():
        if key == "cell_field_param":
            continue  # the oracle's multigroup flattener leaves temperatures at 0 (unused by multigroup physics)
        if key == "n_groups":
            assert mine[key] == value
        else:
            assert np.array_equal(np.array(mine[key], dtype=np.asarray(value).dtype), np.asarray(value)), key


@pytest.mark.parametrize("name", list(decks.DECKS))
def test_flatten_matches_reference_dump(name):
    """tests/golden/<deck>.world.json was written by the reference's own objects (oracle/ref_harness dump)."""
    ref = util.load_golden(f"{name}.world.json")
    mine = _arrays(capi.Driver(text=util.deck_text(name, "surface")).world_json())
    assert [float.fromhex(p) for s in ref["surfaces"] for p in _sphere_params(s)] == _used_params(mine, ref)
    assert mine["cell_material"] == [c["material"] for c in ref["cells"]]
    begin = mine["cell_surface_begin"]
    for i, cell in enumerate(ref["cells"]):
        pairs = sorted(zip(mine["cell_surface_index"][begin[i]:begin[i + 1]], mine["cell_surface_sense"][begin[i]:begin[i + 1]]))
        assert pairs == sorted((s, w) for s, w in cell["surfaces"])  # per-cell order is pointer order in the reference (Q1)
        assert mine["cell_field_param"][6 * i + 4] == float.fromhex(cell["temperature_upper"])
    assert mine["material_aden"] == [float.fromhex(m["aden"]) for m in ref["materials"]]
    mb = mine["material_nuclide_begin"]
    for i, m in enumerate(ref["materials"]):
        pairs = sorted(zip(mine["material_nuclide_index"][mb[i]:mb[i + 1]], mine["material_nuclide_afrac"][mb[i]:mb[i + 1]]))
        assert pairs == sorted((n, float.fromhex(a)) for n, a in m["afracs"])
    G = mine["n_groups"]
    for i, n in enumerate(ref["nuclides"]):
        assert mine["mg_total"][G * i:G * (i + 1)] == [float.fromhex(x) for x in n["total"]]
        for reaction, key in (("capture", "mg_capture"), ("scatter", "mg_scatter"), ("fission", "mg_fission")):
            if reaction in n["reactions"]:
                assert mine[key][G * i:G * (i + 1)] == [float.fromhex(x) for x in n["reactions"][reaction]]
        if n["scatter_probs"]:
            assert mine["mg_scatter_probs"][G * G * i:G * G * (i + 1)] == [float.fromhex(x) for x in n["scatter_probs"]]
        if n["chi"]:
            assert mine["mg_chi"][G * G * i:G * G * (i + 1)] == [float.fromhex(x) for x in n["chi"]]
        if n["nubar"]:
            assert mine["mg_nubar"][G * i:G * (i + 1)] == [float.fromhex(x) for x in n["nubar"]]


def _sphere_params(s):
    n = {"sphere": 4, "planex": 1, "cylinderx": 1}[s["type"]]
    return s["params"][:n]


def _used_params(mine, ref):
    out = []
    for i, s in enumerate(ref["surfaces"]):
        n = {"sphere": 4, "planex": 1, "cylinderx": 1}[s["type"]]
        out += mine["surface_param"][4 * i:4 * i + n]
        assert mine["surface_type"][i] == {"sphere": 0, "planex": 1, "cylinderx": 2}[s["type"]]
    return out


@pytest.mark.parametrize("name,tracking", util.DECK_CASES)
def test_out_text_layout_matches_reference(name, tracking):
    """Before Solve() every score is 0: the text must equal the reference's .out with its numbers zeroed, which
    pins headings, bin edges (std::scientific), separators and blank lines (Estimator.cpp:48-57, Bins.cpp)."""
    import re
    golden = (util.GOLDEN / f"{name}__{tracking}.out").read_text()
    mine = capi.Driver(text=util.deck_text(name, tracking)).output()

    def zero_scores(text):
        out, zero = [], False
        for line in text.split("\n"):
            if line in ("mean", "std dev"):
                zero = True
            elif zero and line and not set(line) <= {"-"}:
                line = re.sub(r"[0-9]\.[0-9]{6}e[+-][0-9]{2}", "0.000000e+00", line)
                zero = False
            out.append(line)
        return "\n".join(out)

    assert mine == zero_scores(golden)


# ---- construction errors: the reference's messages (test_Nuclide.cpp:14-41, test_Material.cpp, test_CSGSurface.cpp)
def _broken(name, old, new):
    text = util.deck_text(name, "surface")
    assert old in text
    return text.replace(old, new, 1)


def _error(text):
    with pytest.raises(capi.MinimcError) as e:
        capi.Driver(text=text)
    return e.value.message


def test_nonexistent_nuclide_message():
    msg = _error(_broken("three_shells", '<nuclide name="oxygen" afrac="0.33"/>', '<nuclide name="nonexistent" afrac="0.33"/>'))
    assert msg.startswith('Nuclide node "nonexistent" not found. Must be one of: ["hydrogen", "oxygen", ')
    assert msg.endswith('", ]')


def test_malformed_multigroup_data_messages():
    text = util.deck_text("critical", "surface")
    import re
    m = re.search(r"<capture>([^<]*)</capture>", text)
    msg = _error(text.replace(m.group(0), "<capture>" + m.group(1) + " 7 8</capture>", 1))
    assert msg == "/minimc/nuclides/multigroup/nuclide/neutron/capture: Expected 1 entries but got 3"
    m = re.search(r"<scatter>([^<]*)</scatter>", text, re.S)
    msg = _error(text.replace(m.group(0), "<scatter>" + m.group(1) + " 0.5</scatter>", 1))
    assert msg == "/minimc/nuclides/multigroup/nuclide/neutron/scatter: Expected 1 entries but got 2"
    msg = _error(text.replace("<neutron>", "<photon>").replace("</neutron>", "</photon>"))
    assert msg == '/minimc/nuclides/multigroup/nuclide: "neutron" node not found'


def test_missing_material_and_surface_messages():
    msg = _error(_broken("leakage_sphere", 'material="absorber"', 'material="nope"'))
    assert msg == 'Material node "nope" not found. Must be one of: ["absorber", ]'
    text = util.deck_text("leakage_sphere", "surface")
    msg = _error(text.replace('<surface name="sphere" sense="-1"/>', '<surface name="ghost" sense="-1"/>', 1))
    assert msg == 'Surface node "ghost" not found. Must be one of: ["sphere", ]'
    msg = _error(text.replace('surface="sphere"', 'surface="ghost"', 1))
    assert msg == 'Surface "ghost" not found. Must be one of: ["sphere", ]'


def test_bins_errors():
    text = util.deck_text("three_shells", "surface")
    assert "<boundaries>1.5</boundaries>" in text
    msg = _error(text.replace("<boundaries>1.5</boundaries>", "<boundaries>1.5 1.5</boundaries>", 1))
    assert msg.endswith("boundaries: nonincreasing elements found: 1.500000 1.500000")
    msg = _error(text.replace('min="-1" max="1"', 'min="1" max="1"', 1))
    assert msg.endswith("linspace: max must be strictly greater than min")


def test_surface_tracking_rejects_nonconstant_temperature():
    text = util.deck_text("leakage_sphere", "surface").replace(
        "</minimc>", '<temperature><linear><bounds lower="300" upper="600"/><intercept b="300"/>'
        '<gradient x="1" y="0" z="0"/></linear></temperature></minimc>')
    assert _error(text) == "Surface tracking with continuous global temperature not allowed"
    # cell delta tracking accepts it (TransportMethod.cpp:39-41)
    capi.Driver(text=util.deck_text("leakage_sphere", "delta").replace(
        "</minimc>", '<temperature><linear><bounds lower="300" upper="600"/><intercept b="300"/>'
        '<gradient x="1" y="0" z="0"/></linear></temperature></minimc>'))


def test_xml_parser_edge_cases():
    text = util.deck_text("leakage_sphere", "surface")
    decorated = '<?xml version="1.0"?>\n<!-- a comment -->\n' + text.replace(
        "<general>", "<general><!-- inner comment -->", 1).replace('name="absorber"', "name='absorber'")
    a, b = capi.Driver(text=text), capi.Driver(text=decorated)
    assert a.world_json() == b.world_json() and a.output() == b.output()
    assert "XML parse error at line" in _error(text.replace("</general>", "</generel>", 1))
    assert "XML parse error" in _error(text + "<extra/>")
    with pytest.raises(capi.MinimcError) as e:
        capi.Driver("/nonexistent/deck.xml")
    assert "File was not found" in e.value.message


def test_missing_table_file_message(tmp_path):
    """continuous_invalid_nuclide_data_path.xml of the reference: 'File not found: <path>' (HDF5DataSet.hpp:100-103)."""
    from minimc_b200 import ce_decks
    text = ce_decks.slab_deck(tmp_path / "no_such_tables")
    msg = _error(text)
    assert msg.startswith("File not found: ") and "no_such_tables" in msg


def test_solve_without_gpu_fails_loudly():
    if capi.load().mmc_device_count() > 0:
        pytest.skip("a GPU is present")
    d = capi.Driver(text=util.deck_text("leakage_sphere", "surface"))
    with pytest.raises(capi.MinimcError) as e:
        d.solve()
    assert "no CUDA device" in e.value.message


def test_perturbations_and_sensitivities_parse_with_the_reference_messages():
    """Perturbation.cpp:101-119, World.cpp:64-81: unknown perturbation / nuclide names are reported as the reference
    reports them; a valid deck builds an EstimatorSet whose sensitivities are named estimator::perturbation."""
    from minimc_b200 import decks
    text = decks.sensitivity_shells(histories=100)
    drv = capi.Driver(text=text)
    assert drv.total_bins == 2 + 2 + 12
    with pytest.raises(capi.MinimcError) as e:
        capi.Driver(text=text.replace('<perturbation name="o-total"/>', '<perturbation name="nope"/>', 1))
    assert 'Perturbation "nope" not found. Must be one of: ["h-total", "o-total", ]' in str(e.value)
    with pytest.raises(capi.MinimcError) as e:
        capi.Driver(text=text.replace('nuclide="oxygen"', 'nuclide="xenon"'))
    assert 'Nuclide "xenon" not found. Must be one of: [' in str(e.value)
