"""The reference's own unit tests for the pieces either side of the hot path, restated against this repo's host code:
the DataTree tokenizer / parser (src/parse/data_tree.rs:568-742), the number readers (src/parse/basics.rs:41-116),
quick_select (src/algorithm.rs:307-351) and log2_64 (src/math.rs:164-188).  Same inputs, same expected values; where the
reference's test looks at a token stream, the assertion here is on the tree those tokens build."""
import ctypes as C

import pytest

from psychopath_b200 import api


def dump(text):
    lib = api.host_lib()
    n = lib.psyh_test_datatree_dump(text.encode(), None, 0)
    if n < 0:
        raise api.PsyError(lib.psyh_last_error().decode())
    buf = C.create_string_buffer(n + 1)
    lib.psyh_test_datatree_dump(text.encode(), buf, n + 1)
    return buf.value.decode()


def floats(text):
    out = (C.c_float * 16)()
    rest = C.c_size_t()
    n = api.host_lib().psyh_test_parse_floats(text.encode(), out, 16, C.byref(rest))
    return list(out[:n]), text[rest.value:]


def uints(text):
    out = (C.c_uint32 * 16)()
    rest = C.c_size_t()
    n = api.host_lib().psyh_test_parse_uints(text.encode(), out, 16, C.byref(rest))
    return list(out[:n]), text[rest.value:]


# ---- data_tree.rs: tokenize_1 .. tokenize_7, parse_1, iter_1 .. iter_3

def test_type_name_after_whitespace_and_comment():
    assert dump("Thing {}") == "Thing{}"                                              # tokenize_1
    assert dump("  \n# gdfgdf gfdg dggdf\\sg dfgsd \n   Thing {}") == "Thing{}"          # tokenize_2
    assert dump(" Thing { }") == "Thing{}"                                              # tokenize_3


def test_identifiers_keep_the_dollar_and_their_escapes():
    assert dump("T $hi_there {}") == "T $hi_there{}"                                  # tokenize_4
    assert dump("T $hi\\ t\\#he\\[re {}") == "T $hi\\ t\\#he\\[re{}"                    # tokenize_5
    # tokenize_6: "$hi the[re" — the identifier ends at the blank, `the` is a type name, `[` opens a leaf
    assert dump("A { B $hi { the[re] } }") == "A{B $hi{the[re]}}"


def test_comment_inside_a_node_and_an_empty_leaf():
    assert dump("Thing $yar { # A comment\n\tThing2 []\n}") == "Thing $yar{Thing2[]}"  # tokenize_7


def test_parse_1_and_the_child_iterators():
    assert dump("\n            Thing {}\n        ") == "Thing{}"                        # parse_1
    text = "\n A {}\n B {}\n A []\n A {}\n B {}\n"
    nodes = dump("R {" + text + "}")[2:-1].split(";")
    assert len([n for n in nodes if n.startswith("A")]) == 3                            # iter_1
    assert len([n for n in nodes if n.startswith("A{")]) == 2                           # iter_2
    text = "\n A []\n B {}\n A {}\n A []\n B {}\n"
    nodes = dump("R {" + text + "}")[2:-1].split(";")
    assert len([n for n in nodes if n.startswith("A[")]) == 2                           # iter_3


def test_unbalanced_text_is_an_error():
    for bad in ("Thing {", "Thing [", "Thing $x [1]", "Thing }", "Thing"):
        with pytest.raises(api.PsyError):
            dump(bad)


# ---- basics.rs: ws_u32_1, ws_usize_1, ws_i32_1 (the non-negative part), ws_f32_1 .. ws_f32_4

@pytest.mark.parametrize("text", ["42", "     42", "42   ", "     42"])
def test_ws_u32_and_ws_f32_take_surrounding_blanks(text):
    assert uints(text) == ([42], "")
    assert floats(text) == ([42.0], "")


def test_a_second_number_is_left_for_the_next_reader():
    assert uints("     42   53") == ([42, 53], "")     # ws_u32("     42   53") == ("53", 42), then 53
    assert floats("     42.5   53") == ([42.5, 53.0], "")
    assert floats("     -42.5   53") == ([-42.5, 53.0], "")


@pytest.mark.parametrize("text,want", [("42.5", 42.5), ("     42.5", 42.5), ("42.5   ", 42.5), ("-42.5", -42.5),
                                       ("     -42.5", -42.5), ("-42.5   ", -42.5)])
def test_ws_f32_values(text, want):
    assert floats(text) == ([want], "")


def test_ws_f32_4_what_is_not_a_number():
    assert floats("a1.0") == ([], "a1.0")                  # ws_f32("a1.0").is_err()
    assert floats("0abc") == ([0.0], "abc")                # all_consuming(ws_f32)("0abc").is_err(): text is left
    assert floats("0.abc 1.2") == ([0.0], "abc 1.2")       # tuple((ws_f32, ws_f32))("0.abc 1.2").is_err()
    assert uints("-42") == ([], "-42")                     # ws_u32 takes digits only
    assert floats("0x10")[0] == [0.0] and floats("inf")[0] == [] and floats("nan")[0] == []   # nom 5's float grammar


def test_fields_read_with_all_consuming_reject_leftovers():
    """psy.rs:277-407: Resolution, SamplesPerPixel, Seed, Fov, FocalDistance, ApertureRadius, Transform (and the light
    and shader scalars) must be exactly their numbers."""
    from scenes import cube_text
    good = cube_text(res=(32, 18), spp=1)
    api.Scene.from_psy_string(good)
    import re
    for pat, repl in ((r"Resolution \[[^\]]*\]", "Resolution [32 18 7]"), (r"Resolution \[[^\]]*\]", "Resolution [32 18 x]"),
                      (r"SamplesPerPixel \[[^\]]*\]", "SamplesPerPixel [1 abc]"), (r"Fov \[[^\]]*\]", "Fov [49.1 0abc]"),
                      (r"Fov \[[^\]]*\]", "Fov [49.1 50]"), (r"Transform \[([^\]]*)\]", r"Transform [\1 1.0]")):
        bad = re.sub(pat, repl, good, count=1)
        assert bad != good
        with pytest.raises(api.PsyError):
            api.Scene.from_psy_string(bad)


# ---- algorithm.rs: quick_select_1 .. quick_select_4

@pytest.mark.parametrize("i", [5, 3, 0, 9])
def test_quick_select(i):
    lst = (C.c_int32 * 10)(8, 9, 7, 4, 6, 1, 0, 5, 3, 2)
    api.host_lib().psyh_test_quick_select(lst, 10, i)
    assert lst[i] == i
    assert sorted(lst) == list(range(10))
    assert all(v <= i for v in lst[:i]) and all(v >= i for v in lst[i + 1:])


# ---- math.rs: log2_64_test

def test_log2_64():
    f = api.host_lib().psyh_test_log2_64
    assert f(0) == 0
    for i in range(64):
        assert f(1 << i) == i
    for lo, add in ((8, 227), (16, 56369), (32, 2514124923)):
        for i in range(lo, 64):
            assert f(((1 << i) + add) & 0xFFFFFFFFFFFFFFFF) == i
