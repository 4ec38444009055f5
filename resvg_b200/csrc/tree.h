// tree.h — the host-side image of a usvg::Tree, as the render traversal (render.cpp, filter_exec.cpp) walks it.
//
// resvg::render takes a finished `usvg::Tree` (crates/usvg/src/tree/mod.rs): groups with resolved transforms, paths with
// resolved paints, clip paths / masks / filters attached to their groups, bounding boxes already computed.  usvg stays
// on the host in Rust; the shim serialises the tree ONCE into the flat stream documented in include/resvg_b200.h
// ("RBT1") and this file is its parsed form.  Field names follow usvg's accessors.
#pragma once

#include <stdint.h>

#include <memory>
#include <string>
#include <vector>

#include "raster_host.h"

struct rb_layer;

namespace rbt {

using rbh::Xform;

struct Rect { float x = 0, y = 0, w = 0, h = 0; };          // usvg Rect / NonZeroRect as x, y, width, height
struct IntRect { int32_t x = 0, y = 0; uint32_t w = 0, h = 0; // tiny_skia::IntRect (width, height >= 1)
    int32_t right() const { return (int32_t)((int64_t)x + (int64_t)w); }
    int32_t bottom() const { return (int32_t)((int64_t)y + (int64_t)h); }
    bool operator==(const IntRect &o) const { return x == o.x && y == o.y && w == o.w && h == o.h; }
    bool operator!=(const IntRect &o) const { return !(*this == o); }
};

struct Group;
struct Tree;

struct Stop { float offset; uint8_t r, g, b; float opacity; };

// usvg::Paint
struct Paint {
    int kind = 0; // 0 Color, 1 LinearGradient, 2 RadialGradient, 3 Pattern
    uint8_t r = 0, g = 0, b = 0;
    float x1 = 0, y1 = 0, x2 = 0, y2 = 0;             // linear
    float cx = 0, cy = 0, rr = 0, fx = 0, fy = 0, fr = 0; // radial
    int spread = 0;                                   // usvg::SpreadMethod: 0 Pad, 1 Reflect, 2 Repeat
    Xform ts;                                         // gradient / pattern transform
    std::vector<Stop> stops;
    Rect rect;                                        // pattern.rect()
    std::unique_ptr<Group> root;                      // pattern.root()
};

struct Fill { Paint paint; float opacity = 1; int rule = 0; /* 0 NonZero, 1 EvenOdd */ };
struct Stroke {
    Paint paint;
    float opacity = 1, width = 1, miterlimit = 4;
    int linecap = 0;  // 0 butt, 1 round, 2 square
    int linejoin = 0; // 0 miter, 1 miter-clip, 2 round, 3 bevel
    std::vector<float> dasharray;
    float dashoffset = 0;
};

struct Path {
    std::string id;
    bool visible = true;
    int paint_order = 0; // 0 FillAndStroke, 1 StrokeAndFill
    bool anti_alias = true; // rendering_mode().use_shape_antialiasing()
    bool has_abs_bbox = false;
    Rect abs_layer_bbox; // Node::abs_layer_bounding_box() = abs_stroke_bounding_box for a path
    std::unique_ptr<Fill> fill;
    std::unique_ptr<Stroke> stroke;
    std::vector<uint8_t> verbs;
    std::vector<float> pts;
    float bounds_w = 0, bounds_h = 0; // path.data().bounds() extent (control-point hull)
};

struct Image {
    std::string id;
    bool visible = true;
    int quality = 2; // tiny_skia::FilterQuality the rendering mode maps to (image.rs:180-187)
    bool has_abs_bbox = false;
    Rect abs_layer_bbox;
    int kind = 0; // 0 ImageKind::SVG, 1 raster (decoded on the host: premultiplied RGBA8, image.rs:62-170)
    std::unique_ptr<Tree> tree;
    uint32_t w = 0, h = 0;
    std::vector<uint8_t> pixels;
};

struct ClipPath { Xform ts; std::unique_ptr<ClipPath> clip_path; std::unique_ptr<Group> root; };
struct Mask { Rect rect; int kind = 0; /* 0 Luminance, 1 Alpha */ std::unique_ptr<Mask> mask; std::unique_ptr<Group> root; };

// usvg::filter
struct Input { int kind = 0; /* 0 SourceGraphic, 1 SourceAlpha, 2 Reference */ std::string name; };
struct TransferFn { int type = 0; std::vector<float> values; float slope = 1, intercept = 0, amplitude = 1, exponent = 1, offset = 0; };
struct Light { int kind = 0; float azimuth = 0, elevation = 0, x = 0, y = 0, z = 0, pax = 0, pay = 0, paz = 0, spec_exp = 1; bool has_cone = false; float cone = 0; };
enum PrimKind {
    P_BLEND = 0, P_DROP_SHADOW, P_FLOOD, P_GAUSSIAN_BLUR, P_OFFSET, P_COMPOSITE, P_MERGE, P_TILE, P_IMAGE, P_COMPONENT_TRANSFER,
    P_COLOR_MATRIX, P_CONVOLVE_MATRIX, P_MORPHOLOGY, P_DISPLACEMENT_MAP, P_TURBULENCE, P_DIFFUSE_LIGHTING, P_SPECULAR_LIGHTING
};
struct Primitive {
    Rect rect;
    int color_interpolation = 1; // 0 sRGB, 1 linearRGB
    std::string result;
    int kind = 0;
    Input in1, in2;
    std::vector<Input> inputs;        // feMerge
    int mode = 0;                     // feBlend: usvg::BlendMode; feComposite: operator 0 over,1 in,2 out,3 atop,4 xor,5 arithmetic;
                                      // feMorphology: 0 erode, 1 dilate; feColorMatrix: kind; feConvolveMatrix: edge mode
    float k[4] = {0, 0, 0, 0};        // arithmetic
    float dx = 0, dy = 0, std_x = 0, std_y = 0; // offset / drop shadow / blur
    uint8_t r = 0, g = 0, b = 0;      // flood / drop shadow / lighting colour
    float opacity = 1;
    std::unique_ptr<Group> root;      // feImage
    TransferFn funcs[4];
    std::vector<float> values;        // colour matrix params / convolve kernel
    uint32_t columns = 0, rows = 0, target_x = 0, target_y = 0;
    float divisor = 1, bias = 0;
    bool preserve_alpha = false;
    float rx = 0, ry = 0;             // morphology
    float scale = 0; int x_channel = 3, y_channel = 3; // displacement map
    float bfx = 0, bfy = 0; uint32_t octaves = 1; int32_t seed = 0; bool stitch = false, fractal = false; // turbulence
    float surface_scale = 1, constant = 1, exponent = 1; Light light; // lighting
};
struct Filter { Rect rect; std::vector<Primitive> primitives; };

struct Node {
    int kind = 0; // 0 Group (also usvg::Node::Text, serialised as its flattened() group), 1 Path, 2 Image
    std::unique_ptr<Group> group;
    std::unique_ptr<Path> path;
    std::unique_ptr<Image> image;
};

struct Group {
    std::string id;
    Xform ts;
    float opacity = 1;
    int blend_mode = 0; // usvg::BlendMode, declaration order (render.rs:145-164)
    bool isolate = false;
    Rect layer_bbox, abs_layer_bbox; // layer_bounding_box() / abs_layer_bounding_box()
    std::unique_ptr<ClipPath> clip_path;
    std::unique_ptr<Mask> mask;
    std::vector<Filter> filters;
    std::vector<Node> children;
    bool should_isolate() const // usvg tree/mod.rs:1201-1208
    {
        return isolate || opacity != 1.0f || clip_path || mask || !filters.empty() || blend_mode != 0;
    }
};

struct Tree { float width = 0, height = 0; Group root; };

// Parses an RBT1 stream; returns nullptr (and a message) on a malformed one.
std::unique_ptr<Tree> parse(const void *blob, size_t len, std::string *err);

} // namespace rbt

struct rb_tree { std::unique_ptr<rbt::Tree> t; };
