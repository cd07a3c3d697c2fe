"""CPU: the Go side of the boundary (go/cubez, go/harness).  No Go toolchain exists in this image, so nothing here
compiles Go; instead every `C.cz_*` call of the cgo package is checked against the prototypes of include/cubezcuda.h
(name and argument count), every entry point the header declares must be bound from Go, and no stub may remain.
The harness mains are parsed and executed through oracle/go2cpp.py by the reference-dump tests."""
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GO_FILES = sorted(glob.glob(os.path.join(ROOT, "go", "cubez", "*.go")))


def header_prototypes():
    text = open(os.path.join(ROOT, "include", "cubezcuda.h")).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(?:int|void \*|const char \*|void)\s*\*?\s*(cz_\w+)\s*\(([^;{]*?)\)\s*;", text, flags=re.S):
        name, args = m.group(1), m.group(2).strip()
        protos[name] = 0 if args in ("", "void") else len(split_args(args))
    return protos


def split_args(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return out


def header_types():
    text = open(os.path.join(ROOT, "include", "cubezcuda.h")).read()
    return set(re.findall(r"\b(cz_\w+)\s*;", " ".join(re.findall(r"typedef[^;]*;|}\s*cz_\w+\s*;", text, flags=re.S))))


def go_calls():
    types = header_types()
    calls = []
    for path in GO_FILES:
        src = re.sub(r"//[^\n]*", "", open(path).read())
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)     # also drops the cgo preamble
        for m in re.finditer(r"\bC\.(cz_\w+)\(", src):
            if m.group(1) in types:          # a conversion such as C.cz_real(x), not a call
                continue
            i, depth = m.end(), 1
            while depth:
                depth += {"(": 1, ")": -1}.get(src[i], 0)
                i += 1
            args = src[m.end():i - 1]
            calls.append((os.path.basename(path), m.group(1), 0 if not args.strip() else len(split_args(args))))
    return calls


def test_header_parses():
    protos = header_prototypes()
    assert len(protos) >= 50 and protos["cz_world_step"] == 4 and protos["cz_real_size"] == 0 and protos["cz_narrowphase"] == 9


def test_every_cgo_call_matches_a_prototype_by_name_and_arity():
    protos = header_prototypes()
    calls = go_calls()
    assert len(calls) >= 50
    for f, name, nargs in calls:
        assert name in protos, f"{f}: C.{name} is not declared in include/cubezcuda.h"
        assert nargs == protos[name], f"{f}: C.{name} called with {nargs} arguments, the header declares {protos[name]}"


def test_every_entry_point_of_the_header_is_bound_from_go():
    bound = {name for _, name, _ in go_calls()}
    missing = sorted(set(header_prototypes()) - bound)
    assert not missing, f"declared in include/cubezcuda.h but never called from go/cubez: {missing}"


def test_no_stub_left_in_the_go_package():
    for path in GO_FILES:
        src = open(path).read()
        assert "elided" not in src and "TODO" not in src and "not implemented" not in src.lower(), path
    names = {os.path.basename(p) for p in GO_FILES}
    assert {"cubez.go", "colliders.go", "contact.go", "world.go", "run.go", "real_f64.go", "real_f32.go"} <= names


def test_reference_api_surface_is_present():
    """Every exported identifier of the reference's rigidbody.go / colliders.go / contact.go a caller can name."""
    src = "\n".join(open(p).read() for p in GO_FILES)
    for ident in ["func NewRigidBody() *RigidBody", "func (b *RigidBody) Clone()", "SetMass(", "SetInfiniteMass(", "HasFiniteMass(", "GetMass(",
                  "GetInverseMass(", "GetTransform(", "GetLastFrameAccelleration(", "GetInverseInertiaTensorWorld(", "SetInertiaTensor(",
                  "SetAwake(", "AddVelocity(", "AddRotation(", "ClearAccumulators(", "Integrate(duration m.Real)", "CalculateDerivedData()",
                  "type Collider interface", "func NewCollisionPlane(", "func NewCollisionSphere(", "func NewCollisionCube(",
                  "CheckAgainstHalfSpace(", "CheckAgainstSphere(", "CheckAgainstCube(",
                  "func CheckForCollisions(one Collider, two Collider, existingContacts []*Contact) (bool, []*Contact)",
                  "func NewContact() *Contact", "func ResolveContacts(maxIterations int, contacts []*Contact, duration m.Real)"]:
        assert ident in src, ident
    assert "build cubez_f32" in src and "build !cubez_f32" in src       # the float32 build tag


def test_four_headless_harnesses_exist_and_import_only_the_reference():
    for name in ("cubedrop", "ballistic", "pile", "integrate_bench"):
        src = open(os.path.join(ROOT, "go", "harness", name + "_headless.go")).read()
        assert '"github.com/tbogdala/cubez"' in src and "package main" in src
        assert "libcubezcuda" not in src and "import \"C\"" not in src     # they run the UNMODIFIED reference
