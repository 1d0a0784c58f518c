// script_recognizer.cpp -- see script_recognizer.h. Host-only C++17, no CUDA.
#include "script_recognizer.h"

#include <functional>

#include <cstdio>

#include <cctype>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <memory>

namespace igbh {

static thread_local std::string g_last_error;
const std::string& last_error() { return g_last_error; }
void set_last_error(const std::string& e) { g_last_error = e; }

[[noreturn]] static void fail(const std::string& what) { throw RecognizeError{what}; }

// ============================================================================================== stage body -> lets
static bool ident_char(char c) { return std::isalnum((unsigned char)c) || c == '_'; }

static std::string trim(const std::string& s) {
    size_t a = 0, b = s.size();
    while (a < b && std::isspace((unsigned char)s[a])) ++a;
    while (b > a && std::isspace((unsigned char)s[b - 1])) --b;
    return s.substr(a, b - a);
}

// Body of the LAST definition of `fn <function>(` in the script (the generated stage follows the standard library,
// src/runtime/shader/ScriptCompiler.cpp:36-51).
static std::string function_body(const std::string& script, const std::string& function) {
    // `fn`, blanks, the name as a whole identifier, blanks, `(` -- whatever the layout (an attribute such as #[export] may precede it)
    size_t at = std::string::npos, after = 0;
    for (size_t q = script.rfind(function); q != std::string::npos; q = q == 0 ? std::string::npos : script.rfind(function, q - 1)) {
        size_t e = q + function.size();
        if (e < script.size() && ident_char(script[e])) continue;
        while (e < script.size() && std::isspace((unsigned char)script[e])) ++e;
        if (e >= script.size() || script[e] != '(') continue;
        size_t b = q;
        while (b > 0 && std::isspace((unsigned char)script[b - 1])) --b;
        if (b == q || b < 2 || script.compare(b - 2, 2, "fn") != 0 || (b >= 3 && ident_char(script[b - 3]))) continue;
        at = q; after = e + 1;
        break;
    }
    if (at == std::string::npos) fail("function '" + function + "' not found in the script");
    // skip the parameter list, then the return type up to the opening brace of the body
    size_t p = after;
    int depth = 1;
    while (p < script.size() && depth > 0) { if (script[p] == '(') ++depth; else if (script[p] == ')') --depth; ++p; }
    while (p < script.size() && script[p] != '{') ++p;
    if (p >= script.size()) fail("function '" + function + "' has no body");
    const size_t open = p;
    depth = 0;
    for (; p < script.size(); ++p) {
        const char c = script[p];
        if (c == '/' && p + 1 < script.size() && script[p + 1] == '/') { while (p < script.size() && script[p] != '\n') ++p; continue; }   // comments may hold anything
        if (c == '"') { ++p; while (p < script.size() && script[p] != '"') ++p; continue; }
        if (c == '{') ++depth;
        else if (c == '}') { if (--depth == 0) return script.substr(open + 1, p - open - 1); }
    }
    fail("unbalanced braces in '" + function + "'");
}

// Splits a block body into statements at top-level ';' (strings, (), {}, [] respected; `//` comments dropped).
static std::vector<std::string> statements(const std::string& body) {
    std::vector<std::string> out;
    std::string cur;
    int depth = 0;
    for (size_t i = 0; i < body.size(); ++i) {
        const char c = body[i];
        if (c == '/' && i + 1 < body.size() && body[i + 1] == '/') { while (i < body.size() && body[i] != '\n') ++i; continue; }
        if (c == '"') { cur += c; ++i; while (i < body.size() && body[i] != '"') cur += body[i++]; cur += '"'; continue; }
        if (c == '(' || c == '{' || c == '[') ++depth;
        if (c == ')' || c == '}' || c == ']') --depth;
        if (c == ';' && depth == 0) { const std::string t = trim(cur); if (!t.empty()) out.push_back(t); cur.clear(); continue; }
        cur += c;
    }
    const std::string t = trim(cur);
    if (!t.empty()) out.push_back(t);
    return out;
}

static StageKind kind_of(const std::string& function) {
    static const struct { const char* name; StageKind k; } table[] = {   // src/runtime/Runtime.cpp:631-657,779
        {"ig_ray_generation_shader", StageKind::RayGeneration}, {"ig_miss_shader", StageKind::Miss}, {"ig_hit_shader", StageKind::Hit},
        {"ig_traversal_shader", StageKind::Traversal}, {"ig_callback_shader", StageKind::Callback}, {"ig_advanced_shadow_shader", StageKind::AdvancedShadow},
        {"ig_tonemap_shader", StageKind::Tonemap}, {"ig_imageinfo_shader", StageKind::ImageInfo}, {"ig_pass_main", StageKind::Pass}, {"ig_bake_shader", StageKind::Bake}};
    for (const auto& e : table) if (function == e.name) return e.k;
    fail("unknown stage entry point '" + function + "'");
}

// `load_simple_<kind>_lights(count, offset, device[, shapes])` bound to `e_<class>` (LoaderLight.cpp:171-186): appends the table's entries
static void embedded_entries(const StageDescriptor& d, const std::string& var, const std::string& which, std::vector<std::string>& names) {
    const auto it = d.index.find(var);
    if (it == d.index.end()) fail(which + ": embedded light table '" + var + "' is not bound");
    const std::string& e = d.lets[it->second].expr;
    static const struct { const char* fn; const char* cls; } kinds[] = {{"load_simple_point_lights", "SimplePointLight"}, {"load_simple_spot_lights", "SimpleSpotLight"},
                                                                        {"load_simple_plane_lights", "SimplePlaneLight"}, {"load_simple_area_lights", "SimpleAreaLight"}};
    for (const auto& k : kinds) {
        if (e.rfind(k.fn, 0) != 0) continue;
        const size_t p = e.find('(');
        char* end = nullptr;
        const long count = std::strtol(e.c_str() + p + 1, &end, 10);
        const long offset = std::strtol(end + 1, nullptr, 10);
        if (offset != (long)names.size()) fail(which + ": embedded table '" + var + "' starts at id " + std::to_string(offset) + ", expected " + std::to_string(names.size()));
        for (long i = 0; i < count; ++i) names.push_back(std::string("@") + k.cls + ":" + std::to_string(i));
        return;
    }
    fail(which + ": '" + e.substr(0, 60) + "' is not an embedded light table this device knows");
}

// `LightTable { count = N, get = @|id:i32| { match(id) { 0 => light_a, ... _ => make_null_light(id) } }}` (LoaderLight.cpp:131-148,199-246), with
// embedded fix-table lights in front: `if id < K { e_<class>.get(id - off) } else if ... else { match(id) { K => light_x, ... } }`, or the whole
// table being one embedded class: `let finite_lights = e_<class>;` (LoaderLight.cpp:188-193)
static std::vector<std::string> light_table(const StageDescriptor& d, const std::string& expr, const std::string& which) {
    std::vector<std::string> names;
    const std::string e = trim(expr);
    if (e.rfind("LightTable", 0) != 0) {
        bool ident = !e.empty();
        for (char ch : e) ident = ident && ident_char(ch);
        if (!ident) fail(which + " is neither a LightTable literal nor an embedded table ('" + e.substr(0, 60) + "...')");
        embedded_entries(d, e, which, names);
        return names;
    }
    const size_t cnt = e.find("count");
    if (cnt == std::string::npos) fail(which + ": LightTable without count");
    size_t p = e.find('=', cnt);
    const long count = std::strtol(e.c_str() + p + 1, nullptr, 10);
    const size_t m = e.find("match");
    // embedded classes: every `<var>.get(id - <off>)` in front of the match
    for (size_t g = e.find(".get("); g != std::string::npos && (m == std::string::npos || g < m); g = e.find(".get(", g + 1)) {
        size_t v0 = g;
        while (v0 > 0 && ident_char(e[v0 - 1])) --v0;
        embedded_entries(d, e.substr(v0, g - v0), which, names);
    }
    if (m == std::string::npos) { if (count == (long)names.size()) return names; fail(which + ": LightTable without match"); }
    p = e.find('{', m);
    while (p != std::string::npos) {
        const size_t arrow = e.find("=>", p);
        if (arrow == std::string::npos) break;
        size_t k = arrow;   // key = token before "=>"
        while (k > 0 && std::isspace((unsigned char)e[k - 1])) --k;
        size_t k0 = k;
        while (k0 > 0 && (ident_char(e[k0 - 1]))) --k0;
        const std::string key = e.substr(k0, k - k0);
        size_t v = arrow + 2;
        while (v < e.size() && std::isspace((unsigned char)e[v])) ++v;
        size_t v1 = v;
        while (v1 < e.size() && ident_char(e[v1])) ++v1;
        const std::string val = e.substr(v, v1 - v);
        if (key != "_") {
            if ((long)names.size() != std::strtol(key.c_str(), nullptr, 10)) fail(which + ": LightTable ids are not consecutive");
            names.push_back(val);
        }
        p = v1;
    }
    if ((long)names.size() != count) fail(which + ": LightTable count does not match its entries");
    return names;
}

StageDescriptor* parse_stage(const std::string& script, const std::string& function) {
    auto d = std::make_unique<StageDescriptor>();
    d->kind = kind_of(function);
    d->function = function;
    const std::string body = function_body(script, function);
    for (const std::string& st : statements(body)) {
        if (st.rfind("let ", 0) != 0) continue;   // calls (`device.handle_hit_shader(...)`, `maybe_unused(x)`) carry no parameters
        size_t p = 4;
        while (p < st.size() && std::isspace((unsigned char)st[p])) ++p;
        size_t q = p;
        if (st[p] == '(') { while (q < st.size() && st[q] != ')') ++q; ++q; }   // tuple pattern: kept under its text
        else while (q < st.size() && ident_char(st[q])) ++q;
        const std::string name = st.substr(p, q - p);
        // skip an optional `: Type` up to the '=' of the binding (types contain no '=')
        const size_t eq = st.find('=', q);
        if (eq == std::string::npos) continue;
        d->index[name] = d->lets.size();
        d->lets.push_back(Binding{name, trim(st.substr(eq + 1))});
    }
    auto has = [&](const char* n) { return d->index.count(n) != 0; };
    auto expr = [&](const char* n) -> const std::string& { return d->lets[d->index.at(n)].expr; };
    if (has("infinite_lights")) { d->infinite_lights = light_table(*d, expr("infinite_lights"), "infinite_lights"); d->has_lights = true; }
    if (has("finite_lights")) d->finite_lights = light_table(*d, expr("finite_lights"), "finite_lights");
    d->has_technique = has("technique");
    d->std_aovs = has("full_technique") && expr("full_technique").find("wrap_infobuffer_renderer") != std::string::npos;
    d->has_camera = has("camera");
    if (has("emitter")) d->list_emitter = expr("emitter").rfind("make_list_emitter", 0) == 0;   // RayGenerationShader.cpp:59-60
    if (d->kind == StageKind::Hit) {
        if (!has("shader")) fail("ig_hit_shader without a material shader binding");
        const std::string& sh = expr("shader");   // ShaderUtils.cpp:127-142
        d->emissive = sh.find("make_emissive_material") != std::string::npos;
        if (!d->emissive && sh.find("make_material") == std::string::npos) fail("unrecognised material shader '" + sh + "'");
        const size_t b = sh.find("bsdf_");
        if (b == std::string::npos) fail("material shader without a bsdf binding: '" + sh + "'");
        size_t e = b;
        while (e < sh.size() && ident_char(sh[e])) ++e;
        d->bsdf_binding = sh.substr(b, e - b);
        if (!has(d->bsdf_binding.c_str())) fail("material shader uses undefined '" + d->bsdf_binding + "'");
        if (sh.find("no_medium_interface") == std::string::npos && has("medium_interface") && expr("medium_interface").find("no_medium_interface") == std::string::npos)
            fail("participating media are not supported by this device");
        if (!d->has_technique) fail("ig_hit_shader without a technique");
    }
    if (d->kind == StageKind::RayGeneration && !d->has_camera && !d->list_emitter) fail("ig_ray_generation_shader without camera or list emitter");
    if (d->kind == StageKind::RayGeneration && has("pixel_sampler") && expr("pixel_sampler").find("make_uniform_pixel_sampler") == std::string::npos)
        fail("pixel sampler '" + expr("pixel_sampler") + "' is not supported (uniform only)");
    return d.release();
}

// ============================================================================================== expression evaluator
namespace {

struct Val {
    enum Kind { Num, Vec, Sym, Ctor } kind = Num;
    float f[4] = {0, 0, 0, 0};
    int n = 1;                    // Vec: number of components
    std::string name;             // Sym: the identifier; Ctor: constructor name
    std::vector<Val> args;        // Ctor
    bool known = true;            // built from literals of the script only (Artic's `?x`: known at specialisation time); registry values are not
    static Val num(float x) { Val v; v.kind = Num; v.f[0] = x; return v; }
    static Val vec(float x, float y, float z) { Val v; v.kind = Vec; v.n = 3; v.f[0] = x; v.f[1] = y; v.f[2] = z; return v; }
    static Val sym(const std::string& s) { Val v; v.kind = Sym; v.name = s; return v; }
};

struct Eval {
    const StageDescriptor& st;
    const Registries& reg;
    int depth = 0;
    std::vector<std::map<std::string, Val>> scopes;   // bindings made inside the blocks being evaluated (`{ let (ru, rv) = ...; f(ru, rv) }`)

    // ---- lexer over one expression string
    struct Cursor { const std::string* s; size_t p; };
    static void ws(Cursor& c) { while (c.p < c.s->size() && std::isspace((unsigned char)(*c.s)[c.p])) ++c.p; }
    static bool eat(Cursor& c, const char* tok) {
        ws(c);
        const size_t n = std::strlen(tok);
        if (c.s->compare(c.p, n, tok) == 0) { c.p += n; return true; }
        return false;
    }
    static bool peek(Cursor& c, const char* tok) { ws(c); return c.s->compare(c.p, std::strlen(tok), tok) == 0; }
    static bool at_end(Cursor& c) { ws(c); return c.p >= c.s->size(); }

    static std::string path(Cursor& c) {   // ident (:: ident | . ident)*
        ws(c);
        std::string out;
        for (;;) {
            const size_t b = c.p;
            while (c.p < c.s->size() && ident_char((*c.s)[c.p])) ++c.p;
            if (c.p == b) fail("identifier expected at '" + c.s->substr(b, 30) + "'");
            out += c.s->substr(b, c.p - b);
            if (c.s->compare(c.p, 2, "::") == 0) { out += "::"; c.p += 2; continue; }
            if (c.p + 1 < c.s->size() && (*c.s)[c.p] == '.' && (std::isalpha((unsigned char)(*c.s)[c.p + 1]) || (*c.s)[c.p + 1] == '_')) { out += '.'; ++c.p; continue; }
            break;
        }
        return out;
    }

    static void skip_type(Cursor& c) {   // after ':' or '->' or 'as': a type name, possibly with generics / & prefixes
        ws(c);
        while (c.p < c.s->size() && ((*c.s)[c.p] == '&' || std::isspace((unsigned char)(*c.s)[c.p]))) ++c.p;
        while (c.p < c.s->size() && (ident_char((*c.s)[c.p]) || (*c.s)[c.p] == ':' )) { if ((*c.s)[c.p] == ':' && c.s->compare(c.p, 2, "::") != 0) break; ++c.p; }
    }

    Val expr(Cursor& c) {
        Val l = term(c);
        for (;;) {
            if (peek(c, "->") ) break;
            if (eat(c, "+")) { l = arith('+', l, term(c)); continue; }
            if (peek(c, "-") && !peek(c, "->")) { eat(c, "-"); l = arith('-', l, term(c)); continue; }
            break;
        }
        return l;
    }
    Val term(Cursor& c) {
        Val l = unary(c);
        for (;;) {
            if (eat(c, "*")) { l = arith('*', l, unary(c)); continue; }
            if (peek(c, "/") && !peek(c, "//")) { eat(c, "/"); l = arith('/', l, unary(c)); continue; }
            break;
        }
        return l;
    }
    Val unary(Cursor& c) {
        if (peek(c, "-") && !peek(c, "->")) { eat(c, "-"); return arith('*', Val::num(-1.0f), unary(c)); }
        Val v = primary(c);
        for (;;) {   // `x as f32`, `ctx.{surf=surf2}` (a record update: the record stays what it was to this evaluator), `texture_dx(..).r`
            ws(c);
            if (c.s->compare(c.p, 3, "as ") == 0) { c.p += 3; skip_type(c); continue; }
            if (c.s->compare(c.p, 2, ".{") == 0) {
                int d = 0;
                ++c.p;
                do { const char ch = (*c.s)[c.p]; if (ch == '{') ++d; else if (ch == '}') --d; ++c.p; } while (c.p < c.s->size() && d > 0);
                continue;
            }
            if (c.p + 1 < c.s->size() && (*c.s)[c.p] == '.' && std::isalpha((unsigned char)(*c.s)[c.p + 1])) {   // component of a run-time value: r | g | b | x ...
                ++c.p;
                while (c.p < c.s->size() && ident_char((*c.s)[c.p])) ++c.p;
                continue;
            }
            break;
        }
        return v;
    }

    static Val arith(char op, const Val& a, const Val& b) {
        if (a.kind == Val::Sym || b.kind == Val::Sym || a.kind == Val::Ctor || b.kind == Val::Ctor) {
            // an expression over run-time quantities (e.g. settings.width as f32 / settings.height as f32): stays symbolic
            Val v; v.kind = Val::Sym; v.name = "(" + (a.kind == Val::Sym ? a.name : std::string("?")) + op + (b.kind == Val::Sym ? b.name : std::string("?")) + ")";
            return v;
        }
        auto f = [&](float x, float y) { return op == '+' ? x + y : op == '-' ? x - y : op == '*' ? x * y : x / y; };   // f32 arithmetic, as Artic evaluates it
        if (a.kind == Val::Num && b.kind == Val::Num) { Val v = Val::num(f(a.f[0], b.f[0])); v.known = a.known && b.known; return v; }
        Val v; v.kind = Val::Vec; v.n = a.kind == Val::Vec ? a.n : b.n; v.known = a.known && b.known;
        for (int i = 0; i < v.n; ++i) v.f[i] = f(a.kind == Val::Vec ? a.f[i] : a.f[0], b.kind == Val::Vec ? b.f[i] : b.f[0]);
        return v;
    }

    Val closure(Cursor& c) {   // |params| [-> Type] (block | expr): the value of a closure is the value of its body
        if (!eat(c, "|")) fail("closure expected");
        while (c.p < c.s->size() && (*c.s)[c.p] != '|') ++c.p;
        eat(c, "|");
        if (eat(c, "->")) skip_type(c);
        if (peek(c, "{")) return block(c);
        return expr(c);
    }

    Val block(Cursor& c) {   // { stmt; stmt; expr }: statements other than the trailing expression are ignored
        if (!eat(c, "{")) fail("block expected");
        const size_t b = c.p;
        int d = 1;
        while (c.p < c.s->size() && d > 0) { const char ch = (*c.s)[c.p]; if (ch == '{') ++d; else if (ch == '}') --d; ++c.p; }
        if (d != 0) fail("unbalanced block");
        const std::string body = c.s->substr(b, c.p - 1 - b);
        const std::vector<std::string> sts = statements(body);
        if (sts.empty()) fail("empty block");
        const std::string& last = sts.back();
        if (last.rfind("let ", 0) == 0) fail("block ends in a binding: '" + last + "'");
        // bodies that are real code (the `match` of a light table or of the AOV selector) stay opaque: they are only an
        // error if a descriptor needs their value, which as_num / as_vec then report
        scopes.emplace_back();
        Val out;
        try {
            for (size_t i = 0; i + 1 < sts.size(); ++i) {   // `let x = e;` / `let (a, b) = e;` in front of the value: visible to what follows
                const std::string& stt = sts[i];
                if (stt.rfind("let ", 0) != 0) continue;
                const size_t eq = stt.find('=');
                if (eq == std::string::npos) continue;
                std::string pat = trim(stt.substr(4, eq - 4));
                const Val v = eval_text(trim(stt.substr(eq + 1)));
                if (!pat.empty() && pat[0] == '(') {   // tuple pattern: component k of a vector-valued right-hand side
                    std::vector<std::string> names; std::string cur;
                    for (char ch : pat) { if (ident_char(ch)) cur += ch; else { if (!cur.empty()) names.push_back(cur); cur.clear(); } }
                    if (!cur.empty()) names.push_back(cur);
                    if (v.kind == Val::Vec) for (size_t k = 0; k < names.size() && (int)k < v.n; ++k) { Val comp = Val::num(v.f[k]); comp.known = v.known; scopes.back()[names[k]] = comp; }
                } else {
                    const size_t colon = pat.find(':');
                    if (colon != std::string::npos) pat = trim(pat.substr(0, colon));
                    scopes.back()[pat] = v;
                }
            }
            out = eval_text(last);
        } catch (const RecognizeError&) { out = Val::sym("<code>"); }
        scopes.pop_back();
        return out;
    }

    Val primary(Cursor& c) {
        ws(c);
        if (at_end(c)) fail("unexpected end of expression");
        const char ch = (*c.s)[c.p];
        if (ch == '(') { eat(c, "("); Val v = expr(c); if (!eat(c, ")")) fail("')' expected in '" + *c.s + "'"); return v; }
        if (ch == '@') { ++c.p; ws(c); if (peek(c, "|")) return closure(c); return primary(c); }   // `@|ctx| ...` or `@name(...)` (a PE annotation)
        if (ch == '|') return closure(c);
        if (ch == '{') return block(c);
        if (ch == '"') { const size_t b = ++c.p; while (c.p < c.s->size() && (*c.s)[c.p] != '"') ++c.p; Val v = Val::sym(c.s->substr(b, c.p - b)); ++c.p; v.name = "\"" + v.name; return v; }
        if (std::isdigit((unsigned char)ch) || (ch == '.' && c.p + 1 < c.s->size() && std::isdigit((unsigned char)(*c.s)[c.p + 1]))) {
            char* end = nullptr;
            const float x = std::strtof(c.s->c_str() + c.p, &end);   // literals are f32 (or i32) in the generated code
            c.p = (size_t)(end - c.s->c_str());
            if (c.p < c.s->size() && (*c.s)[c.p] == ':') { ++c.p; skip_type(c); }   // `0:f32`, `8:i32`
            return Val::num(x);
        }
        std::string name = path(c);
        if (name == "match") fail("match expressions are only understood inside LightTable literals");
        ws(c);
        if (peek(c, "(")) {
            eat(c, "(");
            std::vector<Val> args;
            if (!peek(c, ")")) { do { args.push_back(expr(c)); } while (eat(c, ",")); }
            if (!eat(c, ")")) fail("')' expected after the arguments of " + name);
            return call(name, args);
        }
        if (peek(c, "{") && !name.empty() && std::isupper((unsigned char)name[0])) {   // struct literal: Name{ a = x, b = y }
            eat(c, "{");
            Val v; v.kind = Val::Ctor; v.name = name;
            while (!peek(c, "}")) { path(c); if (!eat(c, "=")) fail("'=' expected in " + name + "{...}"); v.args.push_back(expr(c)); if (!eat(c, ",")) break; }
            if (!eat(c, "}")) fail("'}' expected after " + name + "{...");
            return v;
        }
        return identifier(name);
    }

    Val identifier(const std::string& name) {
        if (name == "true") return Val::num(1.0f);
        if (name == "false") return Val::num(0.0f);
        if (name == "flt_pi") return Val::num(3.14159265359f);                     // core/common.art:3-8
        if (name == "flt_inv_pi") return Val::num(0.31830988618379067154f);
        if (name == "flt_eps") return Val::num(1.1920928955e-07f);
        if (name == "flt_max") return Val::num(3.4028234664e+38f);
        if (name == "color_builtins::black") return Val::vec(0, 0, 0);
        if (name == "color_builtins::white") return Val::vec(1, 1, 1);
        for (size_t k = scopes.size(); k-- > 0;) { const auto f = scopes[k].find(name); if (f != scopes[k].end()) return f->second; }
        const auto it = st.index.find(name);
        if (it != st.index.end()) {
            if (++depth > 64) fail("binding '" + name + "' is recursive");
            Val v = eval_text(st.lets[it->second].expr);
            --depth;
            return v;
        }
        return Val::sym(name);   // run-time quantity: settings.width, ctx.surf, device, entities, mat_id, ...
    }

    static float num_arg(const std::string& fn, const std::vector<Val>& a, size_t i) {
        if (i >= a.size() || a[i].kind != Val::Num) fail(fn + ": argument " + std::to_string(i) + " is not a number");
        return a[i].f[0];
    }
    static const Val& vec_arg(const std::string& fn, const std::vector<Val>& a, size_t i) {
        if (i >= a.size() || a[i].kind != Val::Vec) fail(fn + ": argument " + std::to_string(i) + " is not a vector / colour");
        return a[i];
    }
    static std::string key_arg(const std::string& fn, const std::vector<Val>& a) {
        if (a.empty() || a[0].kind != Val::Sym || a[0].name.empty() || a[0].name[0] != '"') fail(fn + ": parameter name expected");
        return a[0].name.substr(1);
    }

    Val call(const std::string& fn, const std::vector<Val>& a) {
        // ---- registry look-ups (src/artic/driver/registry.art; values stored by ShadingTree.cpp:933-981, defaults as written in the script)
        const bool local = fn.rfind("registry::get_local_parameter_", 0) == 0, global = fn.rfind("registry::get_global_parameter_", 0) == 0;
        if (local || global) {
            const IG::ParameterSet* ps = local ? reg.local : reg.global;
            const std::string key = key_arg(fn, a);
            const std::string ty = fn.substr(fn.rfind('_') + 1);
            if (ps) {
                auto dyn = [](Val v) { v.known = false; return v; };
                if (ty == "i32") { const auto it = ps->IntParameters.find(key); if (it != ps->IntParameters.end()) return dyn(Val::num((float)it->second)); }
                else if (ty == "f32") { const auto it = ps->FloatParameters.find(key); if (it != ps->FloatParameters.end()) return dyn(Val::num(it->second)); }
                else if (ty == "vec3") { const auto it = ps->VectorParameters.find(key); if (it != ps->VectorParameters.end()) return dyn(Val::vec(it->second(0), it->second(1), it->second(2))); }
                else if (ty == "color") { const auto it = ps->ColorParameters.find(key); if (it != ps->ColorParameters.end()) return dyn(Val::vec(it->second(0), it->second(1), it->second(2))); }
                else fail("unsupported registry type in " + fn);
            }
            if (a.size() < 2) fail(fn + ": default value expected");
            Val v = a[1]; v.known = false;   // not in the registry: the default written in the script, as the reference does
            return v;
        }
        if (fn == "make_color") { Val v = Val::vec(num_arg(fn, a, 0), num_arg(fn, a, 1), num_arg(fn, a, 2)); v.known = a[0].known && a[1].known && a[2].known; return v; }
        if (fn == "make_gray_color") { const float g = num_arg(fn, a, 0); return Val::vec(g, g, g); }
        if (fn == "make_vec3") return Val::vec(num_arg(fn, a, 0), num_arg(fn, a, 1), num_arg(fn, a, 2));
        if (fn == "make_vec2") { Val v = Val::vec(num_arg(fn, a, 0), num_arg(fn, a, 1), 0); v.n = 2; return v; }
        if (fn == "vec3_expand" || fn == "color_expand") { const float g = num_arg(fn, a, 0); return Val::vec(g, g, g); }
        if (fn == "vec3_to_2") { Val v = vec_arg(fn, a, 0); v.n = 2; return v; }
        if (fn == "color_mulf" || fn == "vec3_mulf") return arith('*', vec_arg(fn, a, 0), Val::num(num_arg(fn, a, 1)));
        if (fn == "color_mul") return arith('*', vec_arg(fn, a, 0), vec_arg(fn, a, 1));
        if (fn == "make_constant_texture") return vec_arg(fn, a, 0);   // texture/constant: the colour itself on this path
        if ((fn == "vec4_to_color" || fn == "color_to_vec4") && a.size() == 1) return a[0];   // the transpiler's wrapping of a texture look-up (Transpiler.cpp:991,1301)
        if (fn == "microfacet::compute_explicit") {   // core/microfacet.art:427-432
            const float r = num_arg(fn, a, 0), an = num_arg(fn, a, 1);
            const float aspect = (a[1].known && an == 0) ? 1.0f : std::sqrt(1 - std::fmin(std::fmax(an, 0.0f), 1.0f) * 0.99f);
            Val v = Val::vec(r / aspect, r * aspect, 0); v.n = 2; v.known = a[0].known && a[1].known;
            return v;
        }
        if (fn == "maybe_unused") return Val::num(0);
        if (fn == "vec3_normalize") {   // core/vector.art; the length in double, rounded once, then three f32 divisions (scene.py normalize_f32)
            Val v = vec_arg(fn, a, 0);
            const float len = (float)std::sqrt((double)v.f[0] * v.f[0] + (double)v.f[1] * v.f[1] + (double)v.f[2] * v.f[2]);
            for (int i = 0; i < 3; ++i) v.f[i] = v.f[i] / len;
            return v;
        }
        if (fn == "math_builtins::cos") { Val v = Val::num((float)std::cos((double)num_arg(fn, a, 0))); v.known = a[0].known; return v; }
        if (fn == "sun_area_from_srad") { const float r = num_arg(fn, a, 0); Val v = Val::num(3.14159265359f * r * r); v.known = a[0].known; return v; }   // light/sun.art:5
        if (fn == "rad") { Val v = Val::num(num_arg(fn, a, 0) / 180 * 3.14159265359f); v.known = a[0].known; return v; }   // core/common.art:20
        if (fn == "spot_from_power") {   // light/spot.art:1-6; the cosines as in resolve_light
            const float cc = (float)std::cos((double)num_arg(fn, a, 1)), cf = (float)std::cos((double)num_arg(fn, a, 2));
            const float factor = 2 * 3.14159265359f * (1 - 0.5f * cf - 0.5f * cc);
            return arith('*', vec_arg(fn, a, 0), Val::num(1 / factor));
        }
        // a let-bound closure applied to run-time arguments (`bsdf_3(ctx)`, `md_3(ctx)`): its body
        const auto it = st.index.find(fn);
        if (it != st.index.end()) {
            if (++depth > 64) fail("binding '" + fn + "' is recursive");
            Val v = eval_text(st.lets[it->second].expr);
            --depth;
            return v;
        }
        Val v; v.kind = Val::Ctor; v.name = fn; v.args = a;
        return v;
    }

    Val eval_text(const std::string& text) {
        Cursor c{&text, 0};
        Val v = expr(c);
        if (!at_end(c)) fail("trailing text '" + text.substr(c.p, 40) + "' in expression '" + text.substr(0, 80) + "'");
        return v;
    }
    Val binding(const std::string& name) {
        const auto it = st.index.find(name);
        if (it == st.index.end()) fail("binding '" + name + "' not found in " + st.function);
        return eval_text(st.lets[it->second].expr);
    }
};

const Val& ctor_arg(const Val& c, size_t i) {
    if (i >= c.args.size()) fail(c.name + ": argument " + std::to_string(i) + " missing");
    return c.args[i];
}
float as_num(const Val& v, const std::string& what) { if (v.kind != Val::Num) fail(what + " is not a compile-time number"); return v.f[0]; }
const Val& as_vec(const Val& v, const std::string& what) { if (v.kind != Val::Vec) fail(what + " is not a compile-time vector / colour"); return v; }
void put3(float* dst, const Val& v) { dst[0] = v.f[0]; dst[1] = v.f[1]; dst[2] = v.f[2]; }

}  // namespace

// ============================================================================================== descriptors
int TextureTable::add(const igb200_texture& t) {
    for (size_t i = 0; i < records.size(); ++i) if (std::memcmp(&records[i], &t, sizeof(t)) == 0) return (int)i;
    records.push_back(t);
    return (int)records.size() - 1;
}

static int enter_image(TextureTable& t, const std::string& key, const std::function<DeviceImage()>& decode) {
    for (size_t i = 0; i < t.image_keys.size(); ++i) if (t.image_keys[i] == key) return (int)i;
    std::shared_ptr<const DeviceImage> im;
    if (t.cache) { const auto it = t.cache->images.find(key); if (it != t.cache->images.end()) im = it->second; }
    if (!im) { im = std::make_shared<const DeviceImage>(decode()); if (t.cache) t.cache->images[key] = im; }
    t.images.push_back(im);
    t.image_keys.push_back(key);
    return (int)t.images.size() - 1;
}
int TextureTable::image(const std::string& path, bool linear) {
    return enter_image(*this, path + (linear ? "|linear" : "|srgb"), [&] { return load_packed_image(path, linear); });
}
std::shared_ptr<const std::vector<float>> TextureTable::buffer(const std::string& path) {
    if (cache) { const auto it = cache->buffers.find(path); if (it != cache->buffers.end()) return it->second; }
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) fail("cannot open the buffer '" + path + "'");
    std::fseek(f, 0, SEEK_END); const long bytes = std::ftell(f); std::fseek(f, 0, SEEK_SET);
    auto data = std::make_shared<std::vector<float>>((size_t)std::max(0L, bytes) / 4);
    const size_t got = data->empty() ? 0 : std::fread(data->data(), 4, data->size(), f);
    std::fclose(f);
    if (got != data->size()) fail("cannot read the buffer '" + path + "'");
    if (cache) cache->buffers[path] = data;
    return data;
}

int TextureTable::float_image(const std::string& path) {
    return enter_image(*this, path + "|float", [&] {
        FloatImage f = load_float_image(path);
        DeviceImage im;
        im.format = IGB200_IMAGE_RGBA32F; im.width = f.width; im.height = f.height; im.floats.swap(f.rgba);
        return im;
    });
}

// LoaderUtils::inlineTransformAs2d: mat3x3_identity() | make_mat3x3(col0, col1, col2) -> rows 0 and 1
static void texture_transform(const Val& tr, float* out6, const std::string& what) {
    if (tr.kind != Val::Ctor) fail(what + ": texture transform is not a matrix");
    if (tr.name == "mat3x3_identity") { out6[0] = 1; out6[4] = 1; }
    else if (tr.name == "make_mat3x3") {
        for (int c = 0; c < 3; ++c) { const Val col = as_vec(ctor_arg(tr, c), what + ": texture transform column"); out6[c] = col.f[0]; out6[3 + c] = col.f[1]; }
    } else fail(what + ": texture transform '" + tr.name + "' is not understood");
}
static int border_of(const Val& b, const std::string& what) {   // pattern/ImagePattern.cpp:33-51, texture/image.art:9-44
    if (b.kind == Val::Ctor && b.name == "make_repeat_border") return IGB200_BORDER_REPEAT;
    if (b.kind == Val::Ctor && b.name == "make_clamp_border") return IGB200_BORDER_CLAMP;
    if (b.kind == Val::Ctor && b.name == "make_mirror_border") return IGB200_BORDER_MIRROR;
    fail(what + ": image border '" + (b.kind == Val::Ctor ? b.name : std::string("?")) + "' is not understood");
}

// A texture value: `make_checkerboard_texture(make_vec2(sx, sy), color0, color1, transform)` (pattern/CheckerBoardPattern.cpp:13-33) or
// `make_image_texture(border, filter, device.load_packed_image_by_id(id, channels, linear), transform)` (pattern/ImagePattern.cpp:15-73): an
// 8-bit file named through the resource map, decoded by image_io. Float images (`device.load_image_by_id`: EXR / HDR) are the runtime's
// business (IG::Image, tinyexr), which this layer does not link -- they are reported, not guessed.
static int texture_of(const Val& t, TextureTable* textures, const std::string& what) {
    if (t.kind != Val::Ctor) fail(what + " is neither a constant nor a texture");
    if (!textures) fail(what + " is textured but no texture table was given");
    igb200_texture rec;
    std::memset(&rec, 0, sizeof(rec));
    if (t.name == "make_checkerboard_texture") {
        rec.type = IGB200_TEX_CHECKERBOARD;
        const Val sc = as_vec(ctor_arg(t, 0), what + ": checkerboard scale");
        rec.p[0] = sc.f[0]; rec.p[1] = sc.f[1];
        put3(rec.p + 2, as_vec(ctor_arg(t, 1), what + ": checkerboard color0"));
        put3(rec.p + 5, as_vec(ctor_arg(t, 2), what + ": checkerboard color1"));
        texture_transform(ctor_arg(t, 3), rec.transform, what);
    } else if (t.name == "make_image_texture") {
        rec.type = IGB200_TEX_IMAGE;
        const Val& border = ctor_arg(t, 0);
        if (border.kind == Val::Ctor && border.name == "make_split_border") { rec.border_u = border_of(ctor_arg(border, 0), what); rec.border_v = border_of(ctor_arg(border, 1), what); }
        else rec.border_u = rec.border_v = border_of(border, what);
        const Val& filter = ctor_arg(t, 1);
        if (filter.kind != Val::Ctor) fail(what + ": image filter is not a constructor");
        if (filter.name == "make_nearest_filter") rec.filter = IGB200_FILTER_NEAREST;
        else if (filter.name == "make_bilinear_filter") rec.filter = IGB200_FILTER_BILINEAR;
        else if (filter.name == "make_bicubic_filter") rec.filter = IGB200_FILTER_BICUBIC;
        else fail(what + ": image filter '" + filter.name + "' is not understood");
        const Val& img = ctor_arg(t, 2);
        if (img.kind != Val::Ctor) fail(what + ": image is not a load call");
        const bool packed = img.name == "device.load_packed_image_by_id";
        if (!packed && img.name != "device.load_image_by_id") fail(what + ": image source '" + img.name + "' is not understood");
        const int id = (int)as_num(ctor_arg(img, 0), what + ": resource id");
        if (!textures->resource_map || id < 0 || (size_t)id >= textures->resource_map->size()) fail(what + ": resource id " + std::to_string(id) + " is not in the scene's resource map");
        const std::string& file = (*textures->resource_map)[(size_t)id];
        // 8-bit files (PNG) stay packed bytes; float files (OpenEXR: environment maps, the sky texture) become RGBA floats -- image_io.h
        rec.image = packed ? textures->image(file, as_num(ctor_arg(img, 2), what + ": linear flag") != 0) : textures->float_image(file);
        texture_transform(ctor_arg(t, 3), rec.transform, what);
    } else fail(what + ": texture constructor '" + t.name + "' is not supported by this device");
    return textures->add(rec);
}
// A colour parameter that may be a texture look-up (ShadingTree::addColor with a texture name -> `tex_<id>(ctx)`)
static void colour_or_texture(const Val& v, float* dst, int32_t* tex, TextureTable* textures, const std::string& what) {
    if (v.kind == Val::Vec) { put3(dst, v); return; }
    if (!tex) fail(what + " is not a compile-time colour");
    *tex = texture_of(v, textures, what);
    dst[0] = dst[1] = dst[2] = 0;
}

igb200_material resolve_material(const StageDescriptor& hit, const Registries& r, TextureTable* textures) {
    if (hit.kind != StageKind::Hit) fail("resolve_material: not a hit stage");
    Eval ev{hit, r};
    Val b = ev.binding(hit.bsdf_binding);
    if (b.kind != Val::Ctor) fail(hit.bsdf_binding + " is not a BSDF constructor");
    igb200_material m;
    std::memset(&m, 0, sizeof(m));
    m.light_id = -1;
    m.tex[0] = m.tex[1] = -1; m.map_tex = -1;   // no textured parameters, no bump / normal map
    // MapBSDF.cpp:17-55: a wrapper around the material's BSDF that replaces its shading frame (bsdf/map.art:56-68). The map's texture is met
    // before the inner BSDF's, as in the generator.
    if (b.name == "make_bumpmap" || b.name == "make_normalmap") {
        const bool bump = b.name == "make_bumpmap";
        m.map_kind = bump ? IGB200_MAP_BUMP : IGB200_MAP_NORMAL;
        const Val* map = &ctor_arg(b, 2);
        if (bump) {   // texture_dx(<map>, ctx).r, texture_dy(<map>, ctx).r
            if (map->kind != Val::Ctor || map->name != "texture_dx") fail("make_bumpmap: texture_dx(map, ctx).r expected as the x derivative");
            const Val& dy = ctor_arg(b, 3);
            if (dy.kind != Val::Ctor || dy.name != "texture_dy") fail("make_bumpmap: texture_dy(map, ctx).r expected as the y derivative");
            map = &ctor_arg(*map, 0);
        }
        m.map_tex = texture_of(*map, textures, bump ? "bump map" : "normal map");
        m.map_strength = as_num(ctor_arg(b, bump ? 4 : 3), "map strength");
        const Val inner = ctor_arg(b, 1);
        if (inner.kind != Val::Ctor || inner.name == "make_bumpmap" || inner.name == "make_normalmap") fail("bump / normal map without a plain inner BSDF");
        b = inner;
    }
    if (hit.emissive) {   // ShaderUtils.cpp:127-136: the light id lives in the stage's LocalRegistry
        if (!r.local || !r.local->IntParameters.count("_light_id")) fail("emissive material without '_light_id' in its local registry");
        m.light_id = r.local->IntParameters.at("_light_id");
    }
    if (b.name == "make_diffuse_bsdf") {            // DiffuseBSDF.cpp:13-27; bsdf/diffuse.art:55-61
        const float rough = as_num(ctor_arg(b, 1), "diffuse roughness");
        if (rough > 1.1920928955e-07f) fail("rough (Oren-Nayar) diffuse BSDFs are not supported by this device");
        m.bsdf = IGB200_BSDF_DIFFUSE;
        colour_or_texture(ctor_arg(b, 2), m.p, &m.tex[0], textures, "diffuse reflectance");
    } else if (b.name == "make_dielectric_bsdf") {  // DielectricBSDF.cpp:13-41; bsdf/dielectric.art:15-37,195
        const Val& md = ctor_arg(b, 5);
        if (md.kind != Val::Ctor || md.name != "microfacet::make_delta_distribution") fail("rough dielectric BSDFs are not supported by this device");
        if (as_num(ctor_arg(b, 6), "dielectric thin flag") != 0) fail("thin dielectric BSDFs are not supported by this device");
        m.bsdf = IGB200_BSDF_DIELECTRIC;
        m.p[0] = as_num(ctor_arg(b, 1), "ext_ior"); m.p[1] = as_num(ctor_arg(b, 2), "int_ior");
        colour_or_texture(ctor_arg(b, 3), m.p + 2, &m.tex[0], textures, "specular_reflectance");
        colour_or_texture(ctor_arg(b, 4), m.p + 5, &m.tex[1], textures, "specular_transmittance");
    } else if (b.name == "make_conductor_bsdf") {   // ConductorBSDF.cpp:13-35; bsdf/conductor.art:2-27,131-141
        // BSDF::setupRoughness (BSDF.cpp:53-98): no roughness property -> make_delta_distribution; else make_vndf_ggx_distribution(face_normal, local,
        // alpha_u, alpha_v) over the explicit pair or over compute_explicit(roughness, anisotropic) (evaluated inside the md block)
        const Val& md = ctor_arg(b, 4);
        if (md.kind != Val::Ctor) fail("conductor: microfacet distribution expected");
        if (md.name == "microfacet::make_vndf_ggx_distribution") {
            m.distribution = IGB200_MICROFACET_VNDF_GGX;
            m.alpha_u = as_num(ctor_arg(md, 2), "conductor alpha_u"); m.alpha_v = as_num(ctor_arg(md, 3), "conductor alpha_v");
        } else if (md.name != "microfacet::make_delta_distribution")
            fail("microfacet distribution '" + md.name + "' is not supported by this device (delta, vndf_ggx)");
        const Val eta = as_vec(ctor_arg(b, 1), "conductor eta");
        const Val kk = as_vec(ctor_arg(b, 2), "conductor k");
        m.bsdf = IGB200_BSDF_CONDUCTOR;
        put3(m.p, eta); put3(m.p + 3, kk);
        colour_or_texture(ctor_arg(b, 3), m.p + 6, &m.tex[0], textures, "specular_reflectance");
        // conductor.art:133-135: `?eta && ?k && is_black_eps(eta, 1e-4) && is_white_eps(k, 1e-4)` -> make_mirror_bsdf
        bool mirror = eta.known && kk.known;
        for (int i = 0; i < 3; ++i) mirror = mirror && std::fabs(eta.f[i]) <= 1e-4f && std::fabs(kk.f[i] - 1) <= 1e-4f;
        m.p[9] = mirror ? 1.0f : 0.0f;
    } else {
        fail("BSDF constructor '" + b.name + "' is not supported by this device");
    }
    return m;
}

// transform of an environment light: make_mat3x3(c0, c1, c2) -> 9 floats, column major
static void env_transform(const Val& m, float* dst) {
    if (m.kind != Val::Ctor || m.name != "make_mat3x3") fail("environment light transform is not make_mat3x3(...)");
    for (int c = 0; c < 3; ++c) put3(dst + 3 * c, as_vec(ctor_arg(m, c), "environment transform column"));
}

static igb200_light resolve_light(Eval& ev, const std::string& binding, TextureTable* textures) {
    const Val l = ev.binding(binding);
    if (l.kind != Val::Ctor) fail(binding + " is not a light constructor");
    igb200_light out;
    std::memset(&out, 0, sizeof(out));
    out.entity_id = -1;
    if (l.name == "make_environment_light" && ctor_arg(l, 3).kind == Val::Ctor) {
        // EnvironmentLight.cpp:94-101: a textured environment without a cdf (`cdf: none`, or a 1 x 1 bake): sampled uniformly, light/env.art:161-167
        out.type = IGB200_LIGHT_ENV_TEX;
        put3(out.p, as_vec(ctor_arg(l, 2), "environment scale"));
        env_transform(ctor_arg(l, 4), out.p + 3);
        const int32_t tex = texture_of(ctor_arg(l, 3), textures, "environment radiance");
        std::memcpy(&out.p[12], &tex, 4);
    } else if (l.name == "make_environment_light_textured") {
        // EnvironmentLight.cpp:62-93 (cdf = conditional) and SkyLight.cpp:62-72: (id, bbox, scale, tex, cdf::make_cdf_2d_from_buffer(
        // device.load_buffer_by_id(<resource>), size_x, size_y), transform); the buffer is [marginal | conditional rows] as CDF::computeForImage wrote it
        out.type = IGB200_LIGHT_ENV_TEXTURED;
        put3(out.p, as_vec(ctor_arg(l, 2), "environment scale"));
        env_transform(ctor_arg(l, 5), out.p + 3);
        const int32_t tex = texture_of(ctor_arg(l, 3), textures, "environment radiance");
        const Val& cdf = ctor_arg(l, 4);
        if (cdf.kind != Val::Ctor || cdf.name != "cdf::make_cdf_2d_from_buffer") fail("environment cdf '" + (cdf.kind == Val::Ctor ? cdf.name : std::string("?")) + "' is not supported by this device (cdf::make_cdf_2d_from_buffer: conditional)");
        const Val& buf = ctor_arg(cdf, 0);
        if (buf.kind != Val::Ctor || buf.name != "device.load_buffer_by_id") fail("environment cdf buffer is not device.load_buffer_by_id(<resource>)");
        const int id = (int)as_num(ctor_arg(buf, 0), "cdf resource id");
        const int32_t sx = (int32_t)as_num(ctor_arg(cdf, 1), "cdf size_x"), sy = (int32_t)as_num(ctor_arg(cdf, 2), "cdf size_y");
        if (!textures->resource_map || id < 0 || (size_t)id >= textures->resource_map->size()) fail("cdf resource id " + std::to_string(id) + " is not in the scene's resource map");
        const std::string& file = (*textures->resource_map)[(size_t)id];
        const size_t words = (size_t)sy + (size_t)sy * (size_t)sx;
        const int32_t first = (int32_t)textures->aux.size();
        const std::shared_ptr<const std::vector<float>> data = textures->buffer(file);
        if (sx < 1 || sy < 1 || data->size() < words) fail("environment cdf '" + file + "' does not hold " + std::to_string(sy) + " + " + std::to_string(sy) + " x " + std::to_string(sx) + " values");
        textures->aux.insert(textures->aux.end(), data->begin(), data->begin() + (long)words);
        const int32_t tail[4] = {tex, first, sx, sy};
        std::memcpy(&out.p[12], tail, sizeof(tail));
    } else if (l.name == "make_environment_light") {       // EnvironmentLight.cpp:103-110; light/env.art:161-164: colour = scale * texture
        const Val c = Eval::arith('*', as_vec(ctor_arg(l, 2), "environment scale"), as_vec(ctor_arg(l, 3), "environment radiance"));
        out.type = IGB200_LIGHT_ENV_CONST;
        put3(out.p, c);
    } else if (l.name == "make_sun_light") {        // SunLight.cpp:28-57; light/sun.art:10-48
        if (as_num(ctor_arg(l, 5), "sun handle_as_delta flag") != 0) fail("a sun light handled as a delta light is not supported");
        out.type = IGB200_LIGHT_SUN;
        put3(out.p, as_vec(ctor_arg(l, 1), "sun direction"));
        out.p[3] = as_num(ctor_arg(l, 3), "cosine of the sun's half angle");
        put3(out.p + 4, as_vec(ctor_arg(l, 4), "sun radiance"));
    } else if (l.name == "make_directional_light") { // DirectionalLight.cpp:25-41; light/directional.art:1-17
        out.type = IGB200_LIGHT_DIRECTIONAL;
        put3(out.p, as_vec(ctor_arg(l, 1), "light direction"));
        put3(out.p + 3, as_vec(ctor_arg(l, 3), "irradiance"));
    } else if (l.name == "make_point_light") {      // PointLight.cpp:44-62
        out.type = IGB200_LIGHT_POINT;
        put3(out.p, as_vec(ctor_arg(l, 1), "point light origin"));
        put3(out.p + 3, as_vec(ctor_arg(l, 2), "point light intensity"));
    } else if (l.name == "make_spot_light") {       // SpotLight.cpp:62-90; light/spot.art:8-44
        out.type = IGB200_LIGHT_SPOT;
        put3(out.p, as_vec(ctor_arg(l, 1), "spot light origin"));
        const Val d = as_vec(ctor_arg(l, 2), "spot light direction");
        put3(out.p + 3, d);
        // cos(cutoff), cos(falloff): per-light constants, evaluated in double and rounded once (ignis_b200/scene.py: spot_cosines)
        out.p[6] = (float)std::cos((double)as_num(ctor_arg(l, 3), "spot cutoff"));
        out.p[7] = (float)std::cos((double)as_num(ctor_arg(l, 4), "spot falloff"));
        put3(out.p + 8, as_vec(ctor_arg(l, 5), "spot light intensity"));
    } else if (l.name == "make_area_light") {       // AreaLight.cpp:115-220
        const Val& ae = ctor_arg(l, 1);
        const Val rad = as_vec(ctor_arg(l, 2), "area light radiance");
        if (ae.kind != Val::Ctor) fail("area light emitter is not a constructor");
        if (ae.name == "make_plane_area_emitter") {
            out.type = IGB200_LIGHT_PLANE_AREA;
            put3(out.p, as_vec(ctor_arg(ae, 0), "plane origin")); put3(out.p + 3, as_vec(ctor_arg(ae, 1), "plane tangent"));
            put3(out.p + 6, as_vec(ctor_arg(ae, 2), "plane bitangent")); put3(out.p + 9, as_vec(ctor_arg(ae, 3), "plane normal"));
            out.p[12] = as_num(ctor_arg(ae, 4), "plane area");
            for (int k = 0; k < 4; ++k) { const Val t = as_vec(ctor_arg(ae, 5 + k), "plane texcoord"); out.p[13 + 2 * k] = t.f[0]; out.p[14 + 2 * k] = t.f[1]; }
            put3(out.p + 21, rad);
        } else if (ae.name == "make_shape_area_emitter_proxy") {
            out.type = IGB200_LIGHT_SHAPE_AREA;
            out.entity_id = (int32_t)as_num(ctor_arg(ae, 0), "area light entity id");
            put3(out.p, rad);
        } else if (ae.name == "make_sphere_area_emitter") {   // AreaLight.cpp:166-190; light/area.art:260-316
            const Val& ent = ctor_arg(ae, 0);
            const Val& sph = ctor_arg(ae, 1);
            if (ent.kind != Val::Ctor || ent.name != "Entity" || sph.kind != Val::Ctor || sph.name != "Sphere") fail("make_sphere_area_emitter: Entity{...}, Sphere{...} expected");
            out.type = IGB200_LIGHT_SPHERE_AREA;
            out.entity_id = (int32_t)as_num(ctor_arg(ent, 0), "sphere light entity id");
            put3(out.p, rad);
            put3(out.p + 3, as_vec(ctor_arg(sph, 0), "sphere origin"));
            const float radius = as_num(ctor_arg(sph, 1), "sphere radius");
            out.p[6] = radius;
            // compute_ellipsoid_area (shapes/sphere.art:21-27) from the columns of the entity's global matrix: a per-light constant,
            // evaluated in double and rounded once (ignis_b200/scene.py: ellipsoid_area does the same)
            const Val& gm = ctor_arg(ent, 3);
            if (gm.kind != Val::Ctor || gm.name != "make_mat3x4") fail("sphere light: global_mat is not make_mat3x4(...)");
            double axis[3];
            for (int k = 0; k < 3; ++k) { const Val c = as_vec(ctor_arg(gm, k), "global matrix column"); axis[k] = 0; for (int i = 0; i < 3; ++i) { const double x = (double)c.f[i] * (double)radius; axis[k] += x * x; } }
            const double P = (double)1.6f;
            out.p[7] = (float)(4 * (double)3.14159265359f * std::pow((std::pow(axis[0] * axis[1], P / 2) + std::pow(axis[0] * axis[2], P / 2) + std::pow(axis[1] * axis[2], P / 2)) / 3, 1 / P));
        } else {
            fail("area emitter '" + ae.name + "' is not supported by this device");
        }
    } else {
        fail("light constructor '" + l.name + "' is not supported by this device");
    }
    return out;
}

// Entry k of an embedded fix-table (`@<class>:<k>`): what load_simple_*_lights builds from it
static igb200_light embedded_light(const std::string& name, const IG::SceneDatabase* db) {
    const size_t colon = name.find(':');
    const std::string cls = name.substr(1, colon - 1);
    const size_t k = (size_t)std::strtoul(name.c_str() + colon + 1, nullptr, 10);
    if (!db) fail("embedded light table '" + cls + "' but no scene database to read it from");
    const auto it = db->FixTables.find(cls);
    if (it == db->FixTables.end()) fail("the scene database has no fix-table '" + cls + "'");
    const std::vector<IG::uint8>& bytes = it->second.data();
    igb200_light out;
    std::memset(&out, 0, sizeof(out));
    out.entity_id = -1;
    if (cls == "SimplePointLight") {   // light/point.art:20-36, PointLight.cpp:70-78: position xyz, pad, intensity rgb, pad
        if ((k + 1) * 32 > bytes.size()) fail("fix-table 'SimplePointLight' has no entry " + std::to_string(k));
        float f[8];
        std::memcpy(f, bytes.data() + k * 32, 32);
        out.type = IGB200_LIGHT_POINT;
        out.p[0] = f[0]; out.p[1] = f[1]; out.p[2] = f[2];
        out.p[3] = f[4]; out.p[4] = f[5]; out.p[5] = f[6];
    } else fail("embedded lights of class '" + cls + "' are not supported by this device (SimplePointLight)");
    return out;
}

void resolve_lights(const StageDescriptor& stage, const Registries& r, std::vector<igb200_light>& infinite, std::vector<igb200_light>& finite, const IG::SceneDatabase* db, TextureTable* textures) {
    if (!stage.has_lights) fail(stage.function + " carries no light tables");
    Eval ev{stage, r};
    infinite.clear(); finite.clear();
    for (const std::string& b : stage.infinite_lights) infinite.push_back(resolve_light(ev, b, textures));
    for (const std::string& b : stage.finite_lights) finite.push_back(!b.empty() && b[0] == '@' ? embedded_light(b, db) : resolve_light(ev, b, textures));
}

igb200_technique resolve_technique(const StageDescriptor& stage, const Registries& r, std::vector<float>& selector_data) {
    if (!stage.has_technique) fail(stage.function + " carries no technique");
    Eval ev{stage, r};
    const Val t = ev.binding("technique");   // PathTechnique.cpp:35-79
    if (t.kind != Val::Ctor || t.name != "make_path_renderer") fail("technique '" + (t.kind == Val::Ctor ? t.name : std::string("?")) + "' is not supported by this device (path only)");
    igb200_technique out;
    std::memset(&out, 0, sizeof(out));
    selector_data.clear();
    // light selector (LoaderLight.cpp:423-452). The cdf and hierarchy selectors read a buffer the loader wrote into its cache directory and
    // named in the script -- `device.load_buffer("<dir>/light_cdf.bin" | "<dir>/light_hierarchy.bin")` -- which is read here the way
    // Device::loadBuffer would (src/device/Device.cpp: a file of raw 32-bit words). With no finite light both degrade to the uniform
    // selector at specialisation time (light_selector.art:80-81), with no light at all the generator never emits them.
    const Val& sel = ctor_arg(t, 2);
    if (sel.kind != Val::Ctor) fail("light selector is not a constructor call");
    auto buffer_of = [&](const Val& v) -> std::string {
        if (v.kind != Val::Ctor || v.name != "device.load_buffer" || v.args.size() != 1 || v.args[0].kind != Val::Sym || v.args[0].name.empty() || v.args[0].name[0] != '"')
            fail("the light selector's buffer is not device.load_buffer(\"file\")");
        return v.args[0].name.substr(1);
    };
    std::string file;
    if (sel.name == "make_uniform_light_selector") out.light_selector = IGB200_SELECTOR_UNIFORM;
    else if (sel.name == "make_hierarchy_light_selector") { out.light_selector = IGB200_SELECTOR_HIERARCHY; file = buffer_of(ctor_arg(sel, 2)); }
    else if (sel.name == "make_cdf_light_selector") {
        out.light_selector = IGB200_SELECTOR_CDF;
        const Val& cdf = ctor_arg(sel, 2);
        if (cdf.kind != Val::Ctor || cdf.name != "cdf::make_cdf_1d_from_buffer") fail("the cdf light selector's sampler is not cdf::make_cdf_1d_from_buffer(...)");
        file = buffer_of(ctor_arg(cdf, 0));
        if (as_num(ctor_arg(cdf, 2), "cdf offset") != 0) fail("cdf::make_cdf_1d_from_buffer with a non-zero offset");
    } else fail("light selector '" + sel.name + "' is not supported by this device (uniform, cdf, hierarchy)");
    if (out.light_selector != IGB200_SELECTOR_UNIFORM) {
        if (stage.finite_lights.empty()) out.light_selector = IGB200_SELECTOR_UNIFORM;
        else {
            FILE* f = std::fopen(file.c_str(), "rb");
            if (!f) fail("cannot open the light selector's buffer '" + file + "'");
            std::fseek(f, 0, SEEK_END); const long bytes = std::ftell(f); std::fseek(f, 0, SEEK_SET);
            selector_data.resize((size_t)std::max(0L, bytes) / 4);
            const size_t got = selector_data.empty() ? 0 : std::fread(selector_data.data(), 4, selector_data.size(), f);
            std::fclose(f);
            if (got != selector_data.size() || selector_data.empty()) fail("cannot read the light selector's buffer '" + file + "'");
        }
    }
    out.max_depth = (int32_t)as_num(ctor_arg(t, 0), "max_depth");
    out.min_depth = (int32_t)as_num(ctor_arg(t, 1), "min_depth");
    out.clamp = as_num(ctor_arg(t, 4), "clamp");
    out.nee = as_num(ctor_arg(t, 5), "nee flag") != 0 ? 1 : 0;
    return out;
}

igb200_camera resolve_camera(const StageDescriptor& raygen, const Registries& r) {
    if (!raygen.has_camera) fail(raygen.function + " carries no camera");
    Eval ev{raygen, r};
    const Val c = ev.binding("camera");      // PerspectiveCamera.cpp:26-67
    if (c.kind != Val::Ctor || c.name != "make_perspective_camera") fail("camera '" + (c.kind == Val::Ctor ? c.name : std::string("?")) + "' is not supported by this device (perspective only)");
    igb200_camera out;
    std::memset(&out, 0, sizeof(out));
    put3(out.eye, as_vec(ctor_arg(c, 0), "camera eye")); put3(out.dir, as_vec(ctor_arg(c, 1), "camera dir")); put3(out.up, as_vec(ctor_arg(c, 2), "camera up"));
    const Val& sc = ctor_arg(c, 3);
    if (sc.kind != Val::Ctor || (sc.name != "compute_scale_from_hfov" && sc.name != "compute_scale_from_vfov")) fail("camera scale is not compute_scale_from_{h,v}fov");
    out.fov_vertical = sc.name == "compute_scale_from_vfov";
    out.fov = as_num(ctor_arg(sc, 0), "camera fov");
    const Val& asp = ctor_arg(sc, 1);
    out.aspect = asp.kind == Val::Num ? asp.f[0] : 0.0f;   // symbolic = settings.width / settings.height
    out.tmin = as_num(ctor_arg(c, 6), "near clip"); out.tmax = as_num(ctor_arg(c, 7), "far clip");
    return out;
}

}  // namespace igbh
