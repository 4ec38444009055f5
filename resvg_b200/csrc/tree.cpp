// tree.cpp — parser of the RBT1 tree stream (format: include/resvg_b200.h) into rbt::Tree.  Pure host code.
#include "tree.h"

#include <math.h>
#include <string.h>

namespace rbt {
namespace {

struct Reader {
    const uint8_t *p, *end;
    bool ok = true;
    int depth = 0;
    const char *why = "truncated stream";

    bool need(size_t n)
    {
        if (!ok || (size_t)(end - p) < n) { ok = false; return false; }
        return true;
    }
    uint32_t u32()
    {
        if (!need(4)) return 0;
        uint32_t v;
        memcpy(&v, p, 4);
        p += 4;
        return v;
    }
    float f32()
    {
        if (!need(4)) return 0;
        float v;
        memcpy(&v, p, 4);
        p += 4;
        return v;
    }
    bool flag() { return u32() != 0; }
    uint32_t bounded(uint32_t max, const char *what)
    {
        uint32_t v = u32();
        if (ok && v > max) { ok = false; why = what; return 0; }
        return v;
    }
    // a count of records that each occupy at least `min_bytes` of the stream: rejects absurd counts before allocating
    uint32_t count(size_t min_bytes)
    {
        uint32_t v = u32();
        if (ok && (size_t)v * min_bytes > (size_t)(end - p)) { ok = false; why = "count exceeds the stream"; return 0; }
        return v;
    }
    void bytes(void *dst, size_t n)
    {
        const size_t padded = (n + 3) & ~(size_t)3;
        if (!need(padded)) return;
        if (n) memcpy(dst, p, n);
        p += padded;
    }
    std::string str()
    {
        uint32_t n = count(1);
        std::string s(n, '\0');
        bytes(n ? &s[0] : nullptr, n);
        return s;
    }
    Xform xf()
    {
        Xform t;
        t.sx = f32(); t.ky = f32(); t.kx = f32(); t.sy = f32(); t.tx = f32(); t.ty = f32();
        return t;
    }
    Rect rect()
    {
        Rect r;
        r.x = f32(); r.y = f32(); r.w = f32(); r.h = f32();
        return r;
    }
    void rgb(uint8_t *r, uint8_t *g, uint8_t *b)
    {
        uint32_t v = u32();
        *r = (uint8_t)(v & 255); *g = (uint8_t)((v >> 8) & 255); *b = (uint8_t)((v >> 16) & 255);
    }
};

constexpr int kMaxDepth = 1024; // usvg limits element nesting itself; this only bounds the parser's recursion

void parse_group(Reader &r, Group &g);

void parse_base(Reader &r, Paint &p)
{
    p.spread = (int)r.bounded(2, "spread method");
    p.ts = r.xf();
    uint32_t n = r.count(12);
    p.stops.resize(n);
    for (uint32_t i = 0; i < n && r.ok; i++) {
        p.stops[i].offset = r.f32();
        r.rgb(&p.stops[i].r, &p.stops[i].g, &p.stops[i].b);
        p.stops[i].opacity = r.f32();
    }
}

void parse_paint(Reader &r, Paint &p)
{
    p.kind = (int)r.bounded(3, "paint kind");
    switch (p.kind) {
    case 0: r.rgb(&p.r, &p.g, &p.b); break;
    case 1:
        p.x1 = r.f32(); p.y1 = r.f32(); p.x2 = r.f32(); p.y2 = r.f32();
        parse_base(r, p);
        break;
    case 2:
        p.cx = r.f32(); p.cy = r.f32(); p.rr = r.f32(); p.fx = r.f32(); p.fy = r.f32(); p.fr = r.f32();
        parse_base(r, p);
        break;
    default:
        p.rect = r.rect();
        p.ts = r.xf();
        p.root.reset(new Group());
        parse_group(r, *p.root);
    }
}

void parse_path(Reader &r, Path &p)
{
    p.id = r.str();
    p.visible = r.flag();
    p.paint_order = (int)r.bounded(1, "paint order");
    p.anti_alias = r.flag();
    p.has_abs_bbox = r.flag();
    p.abs_layer_bbox = r.rect();
    if (r.flag()) {
        p.fill.reset(new Fill());
        parse_paint(r, p.fill->paint);
        p.fill->opacity = r.f32();
        p.fill->rule = (int)r.bounded(1, "fill rule");
    }
    if (r.flag()) {
        p.stroke.reset(new Stroke());
        Stroke &s = *p.stroke;
        parse_paint(r, s.paint);
        s.opacity = r.f32();
        s.width = r.f32();
        s.miterlimit = r.f32();
        s.linecap = (int)r.bounded(2, "line cap");
        s.linejoin = (int)r.bounded(3, "line join");
        uint32_t nd = r.count(4);
        s.dasharray.resize(nd);
        for (uint32_t i = 0; i < nd && r.ok; i++) s.dasharray[i] = r.f32();
        s.dashoffset = r.f32();
    }
    uint32_t nv = r.count(1);
    p.verbs.resize(nv);
    r.bytes(p.verbs.data(), nv);
    uint32_t np = r.count(8);
    p.pts.resize((size_t)np * 2);
    if (r.need((size_t)np * 8)) {
        if (np) memcpy(p.pts.data(), r.p, (size_t)np * 8);
        r.p += (size_t)np * 8;
    }
    if (!r.ok) return;
    // tiny_skia_path::Path invariants (PathBuilder::finish): starts with a move, point count matches the verbs
    size_t need = 0;
    for (uint32_t i = 0; i < nv; i++) {
        uint8_t v = p.verbs[i];
        if (v > 4) { r.ok = false; r.why = "path verb"; return; }
        need += v == 4 ? 0u : (v <= 1 ? 1u : v);
    }
    if (nv == 0 || p.verbs[0] != 0 || need != np) { r.ok = false; r.why = "path verbs / points mismatch"; return; }
    // Path::bounds(): the hull of all points (computed by tiny-skia when the path is built)
    float l = p.pts[0], t = p.pts[1], rr = l, bb = t;
    for (uint32_t i = 1; i < np; i++) {
        l = fminf(l, p.pts[2 * i]); rr = fmaxf(rr, p.pts[2 * i]);
        t = fminf(t, p.pts[2 * i + 1]); bb = fmaxf(bb, p.pts[2 * i + 1]);
    }
    p.bounds_w = rr - l;
    p.bounds_h = bb - t;
}

void parse_tree_body(Reader &r, Tree &t)
{
    t.width = r.f32();
    t.height = r.f32();
    parse_group(r, t.root);
}

void parse_image(Reader &r, Image &im)
{
    im.id = r.str();
    im.visible = r.flag();
    im.quality = (int)r.bounded(2, "filter quality");
    im.has_abs_bbox = r.flag();
    im.abs_layer_bbox = r.rect();
    im.kind = (int)r.bounded(1, "image kind");
    if (im.kind == 0) {
        im.tree.reset(new Tree());
        parse_tree_body(r, *im.tree);
    } else {
        im.w = r.u32();
        im.h = r.u32();
        if (!r.ok) return;
        if (im.w == 0 || im.h == 0 || (uint64_t)im.w * im.h > (1ull << 30)) { r.ok = false; r.why = "raster image size"; return; }
        const size_t n = (size_t)im.w * im.h * 4;
        if (!r.need(n)) return;
        im.pixels.assign(r.p, r.p + n);
        r.p += n;
    }
}

void parse_clip(Reader &r, ClipPath &c)
{
    if (++r.depth > kMaxDepth) { r.ok = false; r.why = "nesting too deep"; return; }
    c.ts = r.xf();
    if (r.flag()) {
        c.clip_path.reset(new ClipPath());
        parse_clip(r, *c.clip_path);
    }
    c.root.reset(new Group());
    parse_group(r, *c.root);
    r.depth--;
}

void parse_mask(Reader &r, Mask &m)
{
    if (++r.depth > kMaxDepth) { r.ok = false; r.why = "nesting too deep"; return; }
    m.rect = r.rect();
    m.kind = (int)r.bounded(1, "mask kind");
    if (r.flag()) {
        m.mask.reset(new Mask());
        parse_mask(r, *m.mask);
    }
    m.root.reset(new Group());
    parse_group(r, *m.root);
    r.depth--;
}

void parse_input(Reader &r, Input &in)
{
    in.kind = (int)r.bounded(2, "filter input kind");
    if (in.kind == 2) in.name = r.str();
}

void parse_light(Reader &r, Light &l)
{
    l.kind = (int)r.bounded(2, "light kind");
    l.azimuth = r.f32(); l.elevation = r.f32();
    l.x = r.f32(); l.y = r.f32(); l.z = r.f32();
    l.pax = r.f32(); l.pay = r.f32(); l.paz = r.f32();
    l.spec_exp = r.f32();
    l.has_cone = r.flag();
    l.cone = r.f32();
}

void parse_primitive(Reader &r, Primitive &p)
{
    p.rect = r.rect();
    p.color_interpolation = (int)r.bounded(1, "color interpolation");
    p.result = r.str();
    p.kind = (int)r.bounded(P_SPECULAR_LIGHTING, "primitive kind");
    switch (p.kind) {
    case P_BLEND:
        p.mode = (int)r.bounded(15, "blend mode");
        parse_input(r, p.in1); parse_input(r, p.in2);
        break;
    case P_DROP_SHADOW:
        parse_input(r, p.in1);
        p.dx = r.f32(); p.dy = r.f32(); p.std_x = r.f32(); p.std_y = r.f32();
        r.rgb(&p.r, &p.g, &p.b);
        p.opacity = r.f32();
        break;
    case P_FLOOD:
        r.rgb(&p.r, &p.g, &p.b);
        p.opacity = r.f32();
        break;
    case P_GAUSSIAN_BLUR:
        parse_input(r, p.in1);
        p.std_x = r.f32(); p.std_y = r.f32();
        break;
    case P_OFFSET:
        parse_input(r, p.in1);
        p.dx = r.f32(); p.dy = r.f32();
        break;
    case P_COMPOSITE:
        p.mode = (int)r.bounded(5, "composite operator");
        for (float &k : p.k) k = r.f32();
        parse_input(r, p.in1); parse_input(r, p.in2);
        break;
    case P_MERGE: {
        uint32_t n = r.count(4);
        p.inputs.resize(n);
        for (uint32_t i = 0; i < n && r.ok; i++) parse_input(r, p.inputs[i]);
        break;
    }
    case P_TILE: parse_input(r, p.in1); break;
    case P_IMAGE:
        p.root.reset(new Group());
        parse_group(r, *p.root);
        break;
    case P_COMPONENT_TRANSFER:
        parse_input(r, p.in1);
        for (TransferFn &f : p.funcs) {
            f.type = (int)r.bounded(4, "transfer function type");
            uint32_t n = r.count(4);
            f.values.resize(n);
            for (uint32_t i = 0; i < n && r.ok; i++) f.values[i] = r.f32();
            f.slope = r.f32(); f.intercept = r.f32(); f.amplitude = r.f32(); f.exponent = r.f32(); f.offset = r.f32();
        }
        break;
    case P_COLOR_MATRIX: {
        parse_input(r, p.in1);
        p.mode = (int)r.bounded(3, "color matrix kind");
        uint32_t n = r.count(4);
        p.values.resize(n);
        for (uint32_t i = 0; i < n && r.ok; i++) p.values[i] = r.f32();
        const uint32_t want = p.mode == 0 ? 20u : (p.mode == 3 ? 0u : 1u);
        if (r.ok && n != want) { r.ok = false; r.why = "color matrix parameter count"; }
        break;
    }
    case P_CONVOLVE_MATRIX: {
        parse_input(r, p.in1);
        p.columns = r.u32(); p.rows = r.u32(); p.target_x = r.u32(); p.target_y = r.u32();
        p.divisor = r.f32(); p.bias = r.f32();
        p.mode = (int)r.bounded(2, "edge mode");
        p.preserve_alpha = r.flag();
        uint32_t n = r.count(4);
        p.values.resize(n);
        for (uint32_t i = 0; i < n && r.ok; i++) p.values[i] = r.f32();
        if (r.ok && (p.columns == 0 || p.rows == 0 || (uint64_t)p.columns * p.rows != n)) { r.ok = false; r.why = "convolve matrix size"; }
        break;
    }
    case P_MORPHOLOGY:
        parse_input(r, p.in1);
        p.mode = (int)r.bounded(1, "morphology operator");
        p.rx = r.f32(); p.ry = r.f32();
        break;
    case P_DISPLACEMENT_MAP:
        parse_input(r, p.in1); parse_input(r, p.in2);
        p.scale = r.f32();
        p.x_channel = (int)r.bounded(3, "channel"); p.y_channel = (int)r.bounded(3, "channel");
        break;
    case P_TURBULENCE:
        p.bfx = r.f32(); p.bfy = r.f32();
        p.octaves = r.u32();
        p.seed = (int32_t)r.u32();
        p.stitch = r.flag(); p.fractal = r.flag();
        break;
    case P_DIFFUSE_LIGHTING:
    case P_SPECULAR_LIGHTING:
        parse_input(r, p.in1);
        p.surface_scale = r.f32(); p.constant = r.f32(); p.exponent = r.f32();
        r.rgb(&p.r, &p.g, &p.b);
        parse_light(r, p.light);
        break;
    }
}

void parse_group(Reader &r, Group &g)
{
    if (++r.depth > kMaxDepth) { r.ok = false; r.why = "nesting too deep"; return; }
    g.id = r.str();
    g.ts = r.xf();
    g.opacity = r.f32();
    g.blend_mode = (int)r.bounded(15, "blend mode");
    g.isolate = r.flag();
    g.layer_bbox = r.rect();
    g.abs_layer_bbox = r.rect();
    if (r.flag()) {
        g.clip_path.reset(new ClipPath());
        parse_clip(r, *g.clip_path);
    }
    if (r.flag()) {
        g.mask.reset(new Mask());
        parse_mask(r, *g.mask);
    }
    uint32_t nf = r.count(20);
    g.filters.resize(nf);
    for (uint32_t i = 0; i < nf && r.ok; i++) {
        g.filters[i].rect = r.rect();
        uint32_t np = r.count(28);
        g.filters[i].primitives.resize(np);
        for (uint32_t k = 0; k < np && r.ok; k++) parse_primitive(r, g.filters[i].primitives[k]);
    }
    uint32_t nc = r.count(8);
    g.children.resize(nc);
    for (uint32_t i = 0; i < nc && r.ok; i++) {
        Node &n = g.children[i];
        n.kind = (int)r.bounded(2, "node kind");
        if (n.kind == 0) { n.group.reset(new Group()); parse_group(r, *n.group); }
        else if (n.kind == 1) { n.path.reset(new Path()); parse_path(r, *n.path); }
        else { n.image.reset(new Image()); parse_image(r, *n.image); }
    }
    r.depth--;
}

} // namespace

std::unique_ptr<Tree> parse(const void *blob, size_t len, std::string *err)
{
    Reader r{(const uint8_t *)blob, (const uint8_t *)blob + len};
    std::unique_ptr<Tree> t(new Tree());
    if (!blob || len < 12 || r.u32() != 0x31544252u) {
        if (err) *err = "not an RBT1 tree stream";
        return nullptr;
    }
    parse_tree_body(r, *t);
    if (r.ok && r.p != r.end) { r.ok = false; r.why = "trailing bytes"; }
    if (!r.ok) {
        if (err) *err = std::string("malformed tree stream: ") + r.why;
        return nullptr;
    }
    return t;
}

} // namespace rbt
