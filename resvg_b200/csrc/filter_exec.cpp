// filter_exec.cpp — the filter-graph executor: crates/resvg/src/filter/mod.rs:214-1147 (Image / FilterResult, apply,
// apply_inner, get_input, the apply_* drivers, transform_light_source, apply_to_canvas, resolve_std_dev,
// scale_coordinates) restated over device-resident layers.  The per-pixel work is the filter kernels of filters.cu; this
// file decides regions, inputs, colour spaces and the order of the calls exactly as the reference does.
#include <math.h>
#include <string.h>

#include "render.h"

namespace rbr {

using rbt::IntRect;
using rbt::Primitive;
using rbt::Rect;

namespace {

enum Err { E_OK = 0, E_INVALID_REGION, E_NO_RESULTS, E_DEVICE };
enum { CS_SRGB = 0, CS_LINEAR = 1 };

using Shared = std::shared_ptr<rb_layer>;

// filter/mod.rs:214-293
struct Image {
    Shared image;      // all images of one filter have the size of the source layer or of the filter region
    IntRect region;    // the part that holds data, in layer coordinates (only feTile looks at it)
    int color_space = CS_SRGB;
};

struct FilterResult { const std::string *name; Image image; };

struct Exec {
    Renderer &r;
    rb_ctx *rb;
    int dev = RB_OK;

    Err device(int st)
    {
        dev = st;
        return E_DEVICE;
    }
    static Shared own(rb_layer *l) { return Shared(l, LayerDeleter()); }

    // Pixmap::try_create (mod.rs:100-102)
    Err try_create(uint32_t w, uint32_t h, Shared *out)
    {
        rb_layer *l = nullptr;
        int st = rb_layer_create(rb, w, h, &l);
        if (st == RB_ERR_INVALID || st == RB_ERR_OOM) return E_INVALID_REGION;
        if (st != RB_OK) return device(st);
        *out = own(l);
        return E_OK;
    }
    Err clone(const rb_layer *src, Shared *out)
    {
        Err e = try_create(rb_layer_width(src), rb_layer_height(src), out);
        if (e != E_OK) return e;
        int st = rb_layer_copy(out->get(), src);
        return st == RB_OK ? E_OK : device(st);
    }
    // Image::take (mod.rs:268-273): the pixmap itself when nobody else holds it, a copy otherwise
    Err take(Image &&img, Shared *out)
    {
        if (img.image.use_count() == 1) { *out = std::move(img.image); return E_OK; }
        return clone(img.image.get(), out);
    }
    static Image from_image(Shared l, int cs) // mod.rs:232-239
    {
        Image im;
        im.region = IntRect{0, 0, rb_layer_width(l.get()), rb_layer_height(l.get())};
        im.image = std::move(l);
        im.color_space = cs;
        return im;
    }
    // mod.rs:241-266
    Err into_color_space(Image &&img, int cs, Image *out)
    {
        if (cs == img.color_space) { *out = std::move(img); return E_OK; }
        const IntRect region = img.region;
        Shared l;
        Err e = take(std::move(img), &l);
        if (e != E_OK) return e;
        int st = cs == CS_SRGB ? rb_layer_into_srgb(l.get()) : rb_layer_into_linear_rgb(l.get());
        if (st != RB_OK) return device(st);
        out->image = std::move(l);
        out->region = region;
        out->color_space = cs;
        return E_OK;
    }

    // mod.rs:523-564
    Err get_input(const rbt::Input &input, const IntRect &region, rb_layer *source, const std::vector<FilterResult> &results, Image *out)
    {
        if (input.kind == 2) {
            for (size_t i = results.size(); i-- > 0;)
                if (*results[i].name == input.name) { *out = results[i].image; return E_OK; }
            // "Technically unreachable": falls back to SourceGraphic
        }
        Shared l;
        Err e = clone(source, &l);
        if (e != E_OK) return e;
        if (input.kind == 1) { // SourceAlpha: RGB := 0, alpha kept
            static const float kAlphaOnly[20] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0};
            int st = rb_filter_color_matrix(l.get(), 0, kAlphaOnly);
            if (st != RB_OK) return device(st);
        }
        out->image = std::move(l);
        out->region = region;
        out->color_space = CS_SRGB;
        return E_OK;
    }

    static bool approx_zero_ulps(float v) // float-cmp approx_eq_ulps(&0.0, 4)
    {
        if (v == 0.0f) return true;
        if (v != v || v < 0.0f) return false;
        uint32_t bits;
        memcpy(&bits, &v, 4);
        return bits <= 4u;
    }
    // mod.rs:1144-1147
    static void scale_coordinates(float x, float y, const Xform &ts, float *ox, float *oy)
    {
        float sx, sy;
        get_scale(ts, &sx, &sy);
        *ox = x * sx;
        *oy = y * sy;
    }
    // mod.rs:1116-1142
    static bool resolve_std_dev(float std_dx, float std_dy, const Xform &ts, double *ox, double *oy, bool *box_blur)
    {
        float sx, sy;
        scale_coordinates(std_dx, std_dy, ts, &sx, &sy);
        if (approx_zero_ulps(sx) && approx_zero_ulps(sy)) return false;
        if (sx < 0.05f) sx = 0.0f; // tiny sigmas would make the IIR blur produce a transparent image
        if (sy < 0.05f) sy = 0.0f;
        *box_blur = sx >= 2.0f || sy >= 2.0f; // BLUR_SIGMA_THRESHOLD
        *ox = (double)sx;
        *oy = (double)sy;
        return true;
    }
    Err blur(rb_layer *l, double sx, double sy, bool box)
    {
        int st = box ? rb_filter_box_blur(l, sx, sy) : rb_filter_iir_blur(l, sx, sy);
        return st == RB_OK ? E_OK : device(st);
    }
    Err draw(rb_layer *dst, const rb_layer *src, int32_t x, int32_t y, int blend = RB_BLEND_SOURCE_OVER)
    {
        int st = rb_draw_layer(dst, src, x, y, 1.0f, blend); // PixmapPaint::default()
        return st == RB_OK ? E_OK : device(st);
    }

    // mod.rs:581-643
    Err apply_drop_shadow(const Primitive &fe, int cs, const Xform &ts, Image &&input, Image *out)
    {
        float dx, dy;
        scale_coordinates(fe.dx, fe.dy, ts, &dx, &dy);
        Shared pixmap;
        Err e = try_create(rb_layer_width(input.image.get()), rb_layer_height(input.image.get()), &pixmap);
        if (e != E_OK) return e;
        Image in_cs;
        if ((e = into_color_space(std::move(input), cs, &in_cs)) != E_OK) return e;
        Shared input_pixmap;
        if ((e = take(std::move(in_cs), &input_pixmap)) != E_OK) return e;
        Shared shadow;
        if ((e = clone(input_pixmap.get(), &shadow)) != E_OK) return e;
        double sx, sy;
        bool box;
        if (resolve_std_dev(fe.std_x, fe.std_y, ts, &sx, &sy, &box))
            if ((e = blur(shadow.get(), sx, sy, box)) != E_OK) return e;
        // flood: every pixel := the flood colour with its opacity scaled by the pixel's alpha
        int st = rb_filter_flood_alpha(shadow.get(), fe.r, fe.g, fe.b, opacity_to_u8(fe.opacity));
        if (st != RB_OK) return device(st);
        st = cs == CS_SRGB ? rb_layer_into_srgb(shadow.get()) : rb_layer_into_linear_rgb(shadow.get());
        if (st != RB_OK) return device(st);
        if ((e = draw(pixmap.get(), shadow.get(), f2i_(dx), f2i_(dy))) != E_OK) return e;
        if ((e = draw(pixmap.get(), input_pixmap.get(), 0, 0)) != E_OK) return e;
        *out = from_image(std::move(pixmap), cs);
        return E_OK;
    }
    static int32_t f2i_(float v) // Rust `as i32`
    {
        if (v != v) return 0;
        if (v >= 2147483648.0f) return INT32_MAX;
        if (v <= -2147483648.0f) return INT32_MIN;
        return (int32_t)v;
    }

    // mod.rs:645-667
    Err apply_blur(const Primitive &fe, int cs, const Xform &ts, Image &&input, Image *out)
    {
        double sx, sy;
        bool box;
        if (!resolve_std_dev(fe.std_x, fe.std_y, ts, &sx, &sy, &box)) { *out = std::move(input); return E_OK; }
        Image in_cs;
        Err e = into_color_space(std::move(input), cs, &in_cs);
        if (e != E_OK) return e;
        Shared pixmap;
        if ((e = take(std::move(in_cs), &pixmap)) != E_OK) return e;
        if ((e = blur(pixmap.get(), sx, sy, box)) != E_OK) return e;
        *out = from_image(std::move(pixmap), cs);
        return E_OK;
    }

    // mod.rs:669-695
    Err apply_offset(const Primitive &fe, const Xform &ts, Image &&input, Image *out)
    {
        float dx, dy;
        scale_coordinates(fe.dx, fe.dy, ts, &dx, &dy);
        if (approx_zero_ulps(dx) && approx_zero_ulps(dy)) { *out = std::move(input); return E_OK; }
        Shared pixmap;
        Err e = try_create(rb_layer_width(input.image.get()), rb_layer_height(input.image.get()), &pixmap);
        if (e != E_OK) return e;
        if ((e = draw(pixmap.get(), input.image.get(), f2i_(dx), f2i_(dy))) != E_OK) return e;
        *out = from_image(std::move(pixmap), input.color_space);
        return E_OK;
    }

    // mod.rs:697-734 (apply_blend) and 736-804 (apply_composite)
    Err apply_blend_or_composite(const Primitive &fe, int cs, const IntRect &region, Image &&input1, Image &&input2, Image *out)
    {
        Image i1, i2;
        Err e;
        if ((e = into_color_space(std::move(input1), cs, &i1)) != E_OK) return e;
        if ((e = into_color_space(std::move(input2), cs, &i2)) != E_OK) return e;
        Shared pixmap;
        if ((e = try_create(region.w, region.h, &pixmap)) != E_OK) return e;
        if (fe.kind == rbt::P_COMPOSITE && fe.mode == 5) {
            int st = rb_filter_composite_arithmetic(pixmap.get(), i1.image.get(), i2.image.get(), fe.k[0], fe.k[1], fe.k[2], fe.k[3]);
            if (st == RB_ERR_INVALID) return E_INVALID_REGION; // inputs of another size than the region
            if (st != RB_OK) return device(st);
            *out = from_image(std::move(pixmap), cs);
            return E_OK;
        }
        if ((e = draw(pixmap.get(), i2.image.get(), 0, 0)) != E_OK) return e;
        int mode;
        if (fe.kind == rbt::P_BLEND) mode = convert_blend_mode(fe.mode);
        else {
            static const int kOps[5] = {RB_BLEND_SOURCE_OVER, RB_BLEND_SOURCE_IN, RB_BLEND_SOURCE_OUT, RB_BLEND_SOURCE_ATOP, RB_BLEND_XOR};
            mode = kOps[fe.mode];
        }
        if ((e = draw(pixmap.get(), i1.image.get(), 0, 0, mode)) != E_OK) return e;
        *out = from_image(std::move(pixmap), cs);
        return E_OK;
    }

    // mod.rs:806-830
    Err apply_merge(const Primitive &fe, int cs, const IntRect &region, rb_layer *source, const std::vector<FilterResult> &results, Image *out)
    {
        Shared pixmap;
        Err e = try_create(region.w, region.h, &pixmap);
        if (e != E_OK) return e;
        for (const rbt::Input &in : fe.inputs) {
            Image input, in_cs;
            if ((e = get_input(in, region, source, results, &input)) != E_OK) return e;
            if ((e = into_color_space(std::move(input), cs, &in_cs)) != E_OK) return e;
            if ((e = draw(pixmap.get(), in_cs.image.get(), 0, 0)) != E_OK) return e;
        }
        *out = from_image(std::move(pixmap), cs);
        return E_OK;
    }

    // mod.rs:832-844
    Err apply_flood(const Primitive &fe, const IntRect &region, Image *out)
    {
        Shared pixmap;
        Err e = try_create(region.w, region.h, &pixmap);
        if (e != E_OK) return e;
        uint8_t c[4];
        premultiplied_u8((float)fe.r / 255.0f, (float)fe.g / 255.0f, (float)fe.b / 255.0f, (float)opacity_to_u8(fe.opacity) / 255.0f, c);
        int st = rb_layer_fill(pixmap.get(), c[0], c[1], c[2], c[3]);
        if (st != RB_OK) return device(st);
        *out = from_image(std::move(pixmap), CS_SRGB);
        return E_OK;
    }

    // mod.rs:846-868
    Err apply_tile(Image &&input, const IntRect &region, Image *out)
    {
        IntRect sub;
        if (!int_rect_from_xywh((int64_t)input.region.x - region.x, (int64_t)input.region.y - region.y, input.region.w, input.region.h, &sub))
            return E_INVALID_REGION; // translate(..).unwrap()
        rb_layer *t = nullptr;
        int st = rb_layer_clone_rect(input.image.get(), sub.x, sub.y, sub.w, sub.h, &t);
        if (st == RB_ERR_INVALID || st == RB_ERR_OOM) return E_INVALID_REGION;
        if (st != RB_OK) return device(st);
        Shared tile = own(t);
        Shared pixmap;
        Err e = try_create(region.w, region.h, &pixmap);
        if (e != E_OK) return e;
        rb_paint paint;
        memset(&paint, 0, sizeof(paint));
        paint.shader = RB_SHADER_PATTERN;
        paint.pattern = tile.get();
        paint.spread = RB_SPREAD_REPEAT;
        paint.quality = RB_QUALITY_BICUBIC;
        paint.opacity = 1.0f;
        paint.ts[0] = paint.ts[3] = 1.0f;
        paint.ts[4] = (float)sub.x;
        paint.ts[5] = (float)sub.y;
        paint.blend_mode = RB_BLEND_SOURCE_OVER;
        paint.anti_alias = 1;
        st = rb_fill_rect(pixmap.get(), 0.0f, 0.0f, (float)region.w, (float)region.h, &paint, nullptr);
        if (st != RB_OK) return device(st);
        st = rb_layer_flush(pixmap.get()); // the tile layer dies with this scope
        if (st != RB_OK) return device(st);
        *out = from_image(std::move(pixmap), CS_SRGB);
        return E_OK;
    }

    // mod.rs:870-897
    Err apply_image(const Primitive &fe, const IntRect &region, const IntRect &subregion, const Xform &ts, Image *out)
    {
        Shared pixmap;
        Err e = try_create(region.w, region.h, &pixmap);
        if (e != E_OK) return e;
        float sx, sy;
        get_scale(ts, &sx, &sy);
        Xform transform;
        transform.sx = sx; transform.sy = sy;
        transform.tx = (float)subregion.x; transform.ty = (float)subregion.y;
        const Context ctx{IntRect{0, 0, region.w, region.h}};
        r.render_nodes(*fe.root, ctx, transform, pixmap.get());
        if (r.status != RB_OK) return device(r.status);
        *out = from_image(std::move(pixmap), CS_SRGB);
        return E_OK;
    }

    // mod.rs:899-929 (component transfer, colour matrix): on demultiplied pixels
    Err apply_pointwise(const Primitive &fe, int cs, Image &&input, Image *out)
    {
        Image in_cs;
        Err e = into_color_space(std::move(input), cs, &in_cs);
        if (e != E_OK) return e;
        Shared pixmap;
        if ((e = take(std::move(in_cs), &pixmap)) != E_OK) return e;
        int st = rb_layer_demultiply_alpha(pixmap.get());
        if (st != RB_OK) return device(st);
        if (fe.kind == rbt::P_COMPONENT_TRANSFER) {
            rb_transfer_fn fn[4];
            for (int i = 0; i < 4; i++) {
                const rbt::TransferFn &f = fe.funcs[i];
                fn[i].type = f.type;
                fn[i].n_values = (int32_t)f.values.size();
                fn[i].values = f.values.empty() ? nullptr : f.values.data();
                fn[i].slope = f.slope; fn[i].intercept = f.intercept;
                fn[i].amplitude = f.amplitude; fn[i].exponent = f.exponent; fn[i].offset = f.offset;
            }
            st = rb_filter_component_transfer(pixmap.get(), fn);
        } else {
            st = rb_filter_color_matrix(pixmap.get(), fe.mode, fe.values.empty() ? nullptr : fe.values.data());
        }
        if (st != RB_OK) return device(st);
        st = rb_layer_multiply_alpha(pixmap.get());
        if (st != RB_OK) return device(st);
        *out = from_image(std::move(pixmap), cs);
        return E_OK;
    }

    // mod.rs:931-945
    Err apply_convolve_matrix(const Primitive &fe, int cs, Image &&input, Image *out)
    {
        Image in_cs;
        Err e = into_color_space(std::move(input), cs, &in_cs);
        if (e != E_OK) return e;
        Shared pixmap;
        if ((e = take(std::move(in_cs), &pixmap)) != E_OK) return e;
        int st = RB_OK;
        if (fe.preserve_alpha) st = rb_layer_demultiply_alpha(pixmap.get());
        if (st == RB_OK)
            st = rb_filter_convolve_matrix(pixmap.get(), fe.values.data(), fe.columns, fe.rows, fe.target_x, fe.target_y, fe.divisor, fe.bias,
                                           fe.mode, fe.preserve_alpha ? 1 : 0);
        if (st != RB_OK) return device(st);
        *out = from_image(std::move(pixmap), cs);
        return E_OK;
    }

    // mod.rs:947-969
    Err apply_morphology(const Primitive &fe, int cs, const Xform &ts, Image &&input, Image *out)
    {
        Image in_cs;
        Err e = into_color_space(std::move(input), cs, &in_cs);
        if (e != E_OK) return e;
        Shared pixmap;
        if ((e = take(std::move(in_cs), &pixmap)) != E_OK) return e;
        float rx, ry;
        scale_coordinates(fe.rx, fe.ry, ts, &rx, &ry);
        int st;
        if (!(rx > 0.0f && ry > 0.0f)) st = rb_layer_fill(pixmap.get(), 0, 0, 0, 0); // pixmap.clear()
        else st = rb_filter_morphology(pixmap.get(), fe.mode, rx, ry);
        if (st != RB_OK) return device(st);
        *out = from_image(std::move(pixmap), cs);
        return E_OK;
    }

    // mod.rs:971-1000
    Err apply_displacement_map(const Primitive &fe, const IntRect &region, int cs, const Xform &ts, Image &&input1, Image &&input2, Image *out)
    {
        Image i1, i2;
        Err e;
        if ((e = into_color_space(std::move(input1), cs, &i1)) != E_OK) return e;
        if ((e = into_color_space(std::move(input2), cs, &i2)) != E_OK) return e;
        Shared pixmap;
        if ((e = try_create(region.w, region.h, &pixmap)) != E_OK) return e;
        float sx, sy;
        scale_coordinates(fe.scale, fe.scale, ts, &sx, &sy);
        int st = rb_filter_displacement_map(pixmap.get(), i1.image.get(), i2.image.get(), fe.x_channel, fe.y_channel, fe.scale, sx, sy);
        if (st == RB_ERR_INVALID) return E_INVALID_REGION;
        if (st != RB_OK) return device(st);
        *out = from_image(std::move(pixmap), cs);
        return E_OK;
    }

    // mod.rs:1002-1033
    Err apply_turbulence(const Primitive &fe, const IntRect &region, int cs, const Xform &ts, Image *out)
    {
        Shared pixmap;
        Err e = try_create(region.w, region.h, &pixmap);
        if (e != E_OK) return e;
        float sx, sy;
        get_scale(ts, &sx, &sy);
        if (!(approx_zero_ulps(sx) || approx_zero_ulps(sy))) {
            int st = rb_filter_turbulence(pixmap.get(), (double)region.x - (double)ts.tx, (double)region.y - (double)ts.ty, (double)sx, (double)sy,
                                          (double)fe.bfx, (double)fe.bfy, fe.octaves, fe.seed, fe.stitch ? 1 : 0, fe.fractal ? 1 : 0);
            if (st == RB_OK) st = rb_layer_multiply_alpha(pixmap.get());
            if (st != RB_OK) return device(st);
        }
        *out = from_image(std::move(pixmap), cs);
        return E_OK;
    }

    // mod.rs:1063-1098
    static rb_light_source transform_light_source(const rbt::Light &src, const IntRect &region, const Xform &ts)
    {
        rb_light_source l;
        memset(&l, 0, sizeof(l));
        l.kind = src.kind;
        l.azimuth = src.azimuth; l.elevation = src.elevation;
        l.x = src.x; l.y = src.y; l.z = src.z;
        l.points_at_x = src.pax; l.points_at_y = src.pay; l.points_at_z = src.paz;
        l.specular_exponent = src.spec_exp;
        l.has_cone = src.has_cone ? 1 : 0;
        l.limiting_cone_angle = src.cone;
        const float kSqrt2 = 1.41421356237309504880168872420969808f;
        if (src.kind == 1) {
            rbh::Pt p{src.x, src.y};
            rbh::map_points(ts, &p, 1);
            l.x = p.x - (float)region.x;
            l.y = p.y - (float)region.y;
            l.z = src.z * sqrtf(ts.sx * ts.sx + ts.sy * ts.sy) / kSqrt2;
        } else if (src.kind == 2) {
            const float sz = sqrtf(ts.sx * ts.sx + ts.sy * ts.sy) / kSqrt2;
            rbh::Pt p{src.x, src.y};
            rbh::map_points(ts, &p, 1);
            l.x = p.x - (float)region.x;
            l.y = p.y - (float)region.y;
            l.z = src.z * sz;
            rbh::Pt q{src.pax, src.pay};
            rbh::map_points(ts, &q, 1);
            l.points_at_x = q.x - (float)region.x;
            l.points_at_y = q.y - (float)region.y;
            l.points_at_z = src.paz * sz;
        }
        return l;
    }

    // mod.rs:1035-1061 (diffuse) and 1063-... (specular): the input is used as it is, in whatever colour space it has
    Err apply_lighting(const Primitive &fe, const IntRect &region, int cs, const Xform &ts, Image &&input, Image *out)
    {
        Shared pixmap;
        Err e = try_create(region.w, region.h, &pixmap);
        if (e != E_OK) return e;
        const rb_light_source light = transform_light_source(fe.light, region, ts);
        int st;
        if (fe.kind == rbt::P_DIFFUSE_LIGHTING)
            st = rb_filter_diffuse_lighting(pixmap.get(), input.image.get(), fe.surface_scale, fe.constant, fe.r, fe.g, fe.b, &light);
        else
            st = rb_filter_specular_lighting(pixmap.get(), input.image.get(), fe.surface_scale, fe.constant, fe.exponent, fe.r, fe.g, fe.b, &light);
        if (st == RB_ERR_INVALID) return E_INVALID_REGION;
        if (st != RB_OK) return device(st);
        *out = from_image(std::move(pixmap), cs);
        return E_OK;
    }

    // Clears everything outside `sub` (layer coordinates): the four fill_rect(.., Clear) of mod.rs:466-497.
    Err clip_to_subregion(rb_layer *pixmap, const IntRect &sub)
    {
        const float w = (float)rb_layer_width(pixmap), h = (float)rb_layer_height(pixmap);
        rb_paint paint;
        memset(&paint, 0, sizeof(paint));
        paint.shader = RB_SHADER_SOLID;
        paint.color[3] = 1.0f; // Color::BLACK
        paint.opacity = 1.0f;
        paint.ts[0] = paint.ts[3] = 1.0f;
        paint.blend_mode = RB_BLEND_CLEAR;
        paint.anti_alias = 1;
        const float rects[4][4] = {{0.0f, 0.0f, w, (float)sub.y}, {0.0f, 0.0f, (float)sub.x, h}, {(float)sub.right(), 0.0f, w, h}, {0.0f, (float)sub.bottom(), w, h}};
        for (const float *q : rects) {
            int st = rb_fill_rect(pixmap, q[0], q[1], q[2], q[3], &paint, nullptr);
            if (st != RB_OK && st != RB_ERR_INVALID) return device(st); // INVALID: Rect::from_xywh refused it (negative extent)
        }
        return E_OK;
    }

    // mod.rs:345-521
    Err apply_inner(const rbt::Filter &filter, const Xform &ts, rb_layer *source, Image *out)
    {
        Rect fr;
        IntRect region;
        if (!rect_transform(filter.rect, ts, true, &fr) || !to_int_rect(fr, &region)) return E_INVALID_REGION;
        // the source layer was clamped to max_bbox by render_group, the filter rect was not
        const IntRect source_rect{0, 0, rb_layer_width(source), rb_layer_height(source)};
        if (!fit_to_rect(region, source_rect, &region)) return E_INVALID_REGION;

        std::vector<FilterResult> results;
        for (const Primitive &primitive : filter.primitives) {
            Rect pr;
            IntRect subregion;
            if (!rect_transform(primitive.rect, ts, true, &pr) || !to_int_rect(pr, &subregion)) return E_INVALID_REGION;
            // feOffset inherits its region from its input
            if (primitive.kind == rbt::P_OFFSET && primitive.in1.kind == 2)
                for (size_t i = results.size(); i-- > 0;)
                    if (*results[i].name == primitive.in1.name) { subregion = results[i].image.region; break; }
            const int cs = primitive.color_interpolation;
            Image result, input1, input2;
            Err e = E_OK;
            auto in1 = [&]() { return get_input(primitive.in1, region, source, results, &input1); };
            auto in2 = [&]() { return get_input(primitive.in2, region, source, results, &input2); };
            switch (primitive.kind) {
            case rbt::P_BLEND:
            case rbt::P_COMPOSITE:
                if ((e = in1()) == E_OK && (e = in2()) == E_OK)
                    e = apply_blend_or_composite(primitive, cs, region, std::move(input1), std::move(input2), &result);
                break;
            case rbt::P_DROP_SHADOW:
                if ((e = in1()) == E_OK) e = apply_drop_shadow(primitive, cs, ts, std::move(input1), &result);
                break;
            case rbt::P_FLOOD: e = apply_flood(primitive, region, &result); break;
            case rbt::P_GAUSSIAN_BLUR:
                if ((e = in1()) == E_OK) e = apply_blur(primitive, cs, ts, std::move(input1), &result);
                break;
            case rbt::P_OFFSET:
                if ((e = in1()) == E_OK) e = apply_offset(primitive, ts, std::move(input1), &result);
                break;
            case rbt::P_MERGE: e = apply_merge(primitive, cs, region, source, results, &result); break;
            case rbt::P_TILE:
                if ((e = in1()) == E_OK) e = apply_tile(std::move(input1), region, &result);
                break;
            case rbt::P_IMAGE: e = apply_image(primitive, region, subregion, ts, &result); break;
            case rbt::P_COMPONENT_TRANSFER:
            case rbt::P_COLOR_MATRIX:
                if ((e = in1()) == E_OK) e = apply_pointwise(primitive, cs, std::move(input1), &result);
                break;
            case rbt::P_CONVOLVE_MATRIX:
                if ((e = in1()) == E_OK) e = apply_convolve_matrix(primitive, cs, std::move(input1), &result);
                break;
            case rbt::P_MORPHOLOGY:
                if ((e = in1()) == E_OK) e = apply_morphology(primitive, cs, ts, std::move(input1), &result);
                break;
            case rbt::P_DISPLACEMENT_MAP:
                if ((e = in1()) == E_OK && (e = in2()) == E_OK)
                    e = apply_displacement_map(primitive, region, cs, ts, std::move(input1), std::move(input2), &result);
                break;
            case rbt::P_TURBULENCE: e = apply_turbulence(primitive, region, cs, ts, &result); break;
            default:
                if ((e = in1()) == E_OK) e = apply_lighting(primitive, region, cs, ts, std::move(input1), &result);
            }
            if (e != E_OK) return e;

            if (region != subregion) {
                // clip the result: feOffset is not clipped ("We do not support clipping on feOffset")
                IntRect subregion2;
                bool ok = primitive.kind == rbt::P_OFFSET
                              ? int_rect_from_xywh(0, 0, region.w, region.h, &subregion2)
                              : int_rect_from_xywh((int64_t)subregion.x - region.x, (int64_t)subregion.y - region.y, subregion.w, subregion.h, &subregion2);
                if (!ok) return E_INVALID_REGION; // .unwrap()
                const int color_space = result.color_space;
                Shared pixmap;
                if ((e = take(std::move(result), &pixmap)) != E_OK) return e;
                if ((e = clip_to_subregion(pixmap.get(), subregion2)) != E_OK) return e;
                result.image = std::move(pixmap);
                result.region = subregion;
                result.color_space = color_space;
            }
            results.push_back(FilterResult{&primitive.result, std::move(result)});
        }
        if (results.empty()) return E_NO_RESULTS;
        *out = std::move(results.back().image);
        return E_OK;
    }

    // mod.rs:1100-1114
    Err apply_to_canvas(Image &&input, rb_layer *pixmap)
    {
        Image in_srgb;
        Err e = into_color_space(std::move(input), CS_SRGB, &in_srgb);
        if (e != E_OK) return e;
        int st = rb_layer_fill(pixmap, 0, 0, 0, 0);
        if (st != RB_OK) return device(st);
        return draw(pixmap, in_srgb.image.get(), 0, 0);
    }
};

} // namespace

// mod.rs:295-343
void Renderer::apply_filter(const rbt::Filter &filter, const Xform &ts, rb_layer *source)
{
    Exec x{*this, rb};
    Image result;
    Err e = x.apply_inner(filter, ts, source, &result);
    if (e == E_OK) e = x.apply_to_canvas(std::move(result), source);
    if (e == E_DEVICE) { fail(x.dev); return; }
    if (e != E_OK) { // "Clear on error"
        int st = rb_layer_fill(source, 0, 0, 0, 0);
        if (st != RB_OK) fail(st);
    }
}

} // namespace rbr
