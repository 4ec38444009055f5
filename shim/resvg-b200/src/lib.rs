//! resvg-b200 — the Rust host shim: `resvg::render(&usvg::Tree, Transform, &mut PixmapMut)` and `resvg::render_node`
//! (crates/resvg/src/lib.rs:34-70) on a B200 through libresvg_b200.so.
//!
//! usvg (parsing, CSS, text shaping, the tree and its bounding boxes) stays exactly where it is; this crate walks the
//! finished tree ONCE, writes it as the flat "RBT1" stream documented in include/resvg_b200.h, and hands it to the
//! library, whose C++ traversal (csrc/render.cpp, csrc/filter_exec.cpp) and CUDA kernels do everything below
//! `resvg::render`.  One call crosses the boundary per document; pixels cross PCIe once, at the end.
//!
//! This source is not compiled in the repository's container (the image has no Rust toolchain).  The stream writer below
//! is mirrored line for line by `resvg_b200/tree.py::serialize`, which IS exercised: every golden scene of the test suite
//! reaches the library through that writer's output.
//!
//! Drop-in use inside resvg (crates/resvg/src/lib.rs):
//!
//! ```ignore
//! pub fn render(tree: &usvg::Tree, transform: tiny_skia::Transform, pixmap: &mut tiny_skia::PixmapMut) {
//!     resvg_b200::with_default_device(|dev| dev.render(tree, transform, pixmap)).expect("B200 renderer");
//! }
//! ```
pub mod ffi;

use std::ffi::{CStr, CString};
use std::ptr;

use tiny_skia::{PixmapMut, Transform};

#[derive(Debug)]
pub enum Error {
    /// No usable CUDA device: the library has no CPU fallback.
    NoDevice,
    /// The library refused the tree stream (a bug in this writer).
    MalformedTree,
    /// `render_node`'s `None`: unknown id or a zero-sized node.
    NoSuchNode,
    Cuda(String),
}

/// One per GPU (rb_ctx).
pub struct Device {
    ctx: *mut ffi::rb_ctx,
}

unsafe impl Send for Device {}

impl Device {
    pub fn new(index: i32) -> Result<Self, Error> {
        let mut ctx = ptr::null_mut();
        match unsafe { ffi::rb_ctx_create(index, &mut ctx) } {
            ffi::RB_OK => Ok(Device { ctx }),
            _ => Err(Error::NoDevice),
        }
    }

    fn check(&self, st: i32) -> Result<(), Error> {
        match st {
            ffi::RB_OK => Ok(()),
            ffi::RB_ERR_INVALID => Err(Error::MalformedTree),
            _ => Err(Error::Cuda(unsafe { CStr::from_ptr(ffi::rb_last_error(self.ctx)) }.to_string_lossy().into_owned())),
        }
    }

    /// A tree, serialised and parsed once; render it as often as needed (`resvg_render_tree` of the C API).
    pub fn upload(&self, tree: &usvg::Tree) -> Result<Tree, Error> {
        let stream = serialize(tree);
        let mut h = ptr::null_mut();
        self.check(unsafe { ffi::rb_tree_parse(stream.as_ptr().cast(), stream.len(), &mut h) })?;
        Ok(Tree { h })
    }

    /// `resvg::render`: draws over the pixmap's current content (premultiplied RGBA8, as tiny-skia keeps it).
    pub fn render(&self, tree: &usvg::Tree, transform: Transform, pixmap: &mut PixmapMut) -> Result<(), Error> {
        let t = self.upload(tree)?;
        self.render_uploaded(&t, transform, pixmap)
    }

    pub fn render_uploaded(&self, tree: &Tree, transform: Transform, pixmap: &mut PixmapMut) -> Result<(), Error> {
        let ts = ts6(transform);
        let (w, h) = (pixmap.width(), pixmap.height());
        self.check(unsafe { ffi::rb_render_to_host(self.ctx, tree.h, ts.as_ptr(), w, h, pixmap.data_mut().as_mut_ptr()) })
    }

    /// `resvg::render_node` (lib.rs:55-70): `Err(NoSuchNode)` is the reference's `None`.
    pub fn render_node(&self, tree: &Tree, id: &str, transform: Transform, pixmap: &mut PixmapMut) -> Result<(), Error> {
        let id = CString::new(id).map_err(|_| Error::NoSuchNode)?;
        let ts = ts6(transform);
        let mut layer = ptr::null_mut();
        self.check(unsafe { ffi::rb_layer_create(self.ctx, pixmap.width(), pixmap.height(), &mut layer) })?;
        let res = (|| {
            self.check(unsafe { ffi::rb_layer_upload(layer, pixmap.data().as_ptr()) })?;
            match unsafe { ffi::rb_render_node(self.ctx, tree.h, id.as_ptr(), ts.as_ptr(), layer) } {
                ffi::RB_ERR_INVALID => return Err(Error::NoSuchNode),
                st => self.check(st)?,
            }
            self.check(unsafe { ffi::rb_layer_download(layer, pixmap.data_mut().as_mut_ptr()) })
        })();
        unsafe { ffi::rb_layer_destroy(layer) };
        res
    }

    /// `node.abs_layer_bounding_box()` as the library sees it: the pixmap size `render_node` expects.
    pub fn node_bbox(&self, tree: &Tree, id: &str) -> Option<[f32; 4]> {
        let id = CString::new(id).ok()?;
        let mut out = [0f32; 4];
        (unsafe { ffi::rb_tree_node_bbox(tree.h, id.as_ptr(), out.as_mut_ptr()) } == ffi::RB_OK).then_some(out)
    }
}

impl Drop for Device {
    fn drop(&mut self) {
        unsafe { ffi::rb_ctx_destroy(self.ctx) }
    }
}

pub struct Tree {
    h: *mut ffi::rb_tree,
}

impl Drop for Tree {
    fn drop(&mut self) {
        unsafe { ffi::rb_tree_destroy(self.h) }
    }
}

/// A process-wide device 0, created on first use.
pub fn with_default_device<R>(f: impl FnOnce(&Device) -> Result<R, Error>) -> Result<R, Error> {
    use std::sync::{Mutex, OnceLock};
    static DEV: OnceLock<Mutex<Option<Device>>> = OnceLock::new();
    let mut guard = DEV.get_or_init(|| Mutex::new(None)).lock().unwrap();
    if guard.is_none() {
        *guard = Some(Device::new(0)?);
    }
    f(guard.as_ref().unwrap())
}

fn ts6(t: Transform) -> [f32; 6] {
    [t.sx, t.ky, t.kx, t.sy, t.tx, t.ty]
}

// ---------------------------------------------------------------------------------------------------------------------
// The RBT1 writer (format: include/resvg_b200.h).  Little-endian 4-byte words.
// ---------------------------------------------------------------------------------------------------------------------
struct W(Vec<u8>);

impl W {
    fn u32(&mut self, v: u32) { self.0.extend_from_slice(&v.to_le_bytes()) }
    fn i32(&mut self, v: i32) { self.0.extend_from_slice(&v.to_le_bytes()) }
    fn f32(&mut self, v: f32) { self.0.extend_from_slice(&v.to_le_bytes()) }
    fn flag(&mut self, v: bool) { self.u32(v as u32) }
    fn raw(&mut self, b: &[u8]) {
        self.0.extend_from_slice(b);
        while self.0.len() % 4 != 0 { self.0.push(0) }
    }
    fn str(&mut self, s: &str) { self.u32(s.len() as u32); self.raw(s.as_bytes()) }
    fn xf(&mut self, t: usvg::Transform) { for v in [t.sx, t.ky, t.kx, t.sy, t.tx, t.ty] { self.f32(v) } }
    fn rect(&mut self, x: f32, y: f32, w: f32, h: f32) { self.f32(x); self.f32(y); self.f32(w); self.f32(h) }
    fn nz(&mut self, r: usvg::NonZeroRect) { self.rect(r.x(), r.y(), r.width(), r.height()) }
    fn rgb(&mut self, c: usvg::Color) { self.u32(c.red as u32 | (c.green as u32) << 8 | (c.blue as u32) << 16) }
}

pub fn serialize(tree: &usvg::Tree) -> Vec<u8> {
    let mut w = W(Vec::with_capacity(1 << 16));
    w.u32(0x3154_4252); // "RBT1"
    write_tree(&mut w, tree);
    w.0
}

fn write_tree(w: &mut W, tree: &usvg::Tree) {
    w.f32(tree.size().width());
    w.f32(tree.size().height());
    write_group(w, tree.root());
}

fn write_group(w: &mut W, g: &usvg::Group) {
    w.str(g.id());
    w.xf(g.transform());
    w.f32(g.opacity().get());
    w.u32(g.blend_mode() as u32); // declaration order = render.rs:145-164
    w.flag(g.isolate());
    w.nz(g.layer_bounding_box());
    w.nz(g.abs_layer_bounding_box());
    w.flag(g.clip_path().is_some());
    if let Some(c) = g.clip_path() { write_clip(w, c) }
    w.flag(g.mask().is_some());
    if let Some(m) = g.mask() { write_mask(w, m) }
    w.u32(g.filters().len() as u32);
    for f in g.filters() {
        w.nz(f.rect());
        w.u32(f.primitives().len() as u32);
        for p in f.primitives() { write_primitive(w, p) }
    }
    w.u32(g.children().len() as u32);
    for n in g.children() {
        match n {
            usvg::Node::Group(g) => { w.u32(0); write_group(w, g) }
            usvg::Node::Text(t) => { w.u32(0); write_group(w, t.flattened()) } // render.rs:43-45
            usvg::Node::Path(p) => { w.u32(1); write_path(w, p) }
            usvg::Node::Image(i) => { w.u32(2); write_image(w, i) }
        }
    }
}

fn write_clip(w: &mut W, c: &usvg::ClipPath) {
    w.xf(c.transform());
    w.flag(c.clip_path().is_some());
    if let Some(n) = c.clip_path() { write_clip(w, n) }
    write_group(w, c.root());
}

fn write_mask(w: &mut W, m: &usvg::Mask) {
    w.nz(m.rect());
    w.u32(matches!(m.kind(), usvg::MaskType::Alpha) as u32);
    w.flag(m.mask().is_some());
    if let Some(n) = m.mask() { write_mask(w, n) }
    write_group(w, m.root());
}

fn write_base(w: &mut W, g: &usvg::BaseGradient) {
    w.u32(match g.spread_method() { usvg::SpreadMethod::Pad => 0, usvg::SpreadMethod::Reflect => 1, usvg::SpreadMethod::Repeat => 2 });
    w.xf(g.transform());
    w.u32(g.stops().len() as u32);
    for s in g.stops() {
        w.f32(s.offset().get());
        w.rgb(s.color());
        w.f32(s.opacity().get());
    }
}

fn write_paint(w: &mut W, p: &usvg::Paint) {
    match p {
        usvg::Paint::Color(c) => { w.u32(0); w.rgb(*c) }
        usvg::Paint::LinearGradient(lg) => {
            w.u32(1);
            for v in [lg.x1(), lg.y1(), lg.x2(), lg.y2()] { w.f32(v) }
            write_base(w, lg);
        }
        usvg::Paint::RadialGradient(rg) => {
            w.u32(2);
            for v in [rg.cx(), rg.cy(), rg.r().get(), rg.fx(), rg.fy(), rg.fr().get()] { w.f32(v) }
            write_base(w, rg);
        }
        usvg::Paint::Pattern(pt) => {
            w.u32(3);
            w.nz(pt.rect());
            w.xf(pt.transform());
            write_group(w, pt.root());
        }
    }
}

fn write_path(w: &mut W, p: &usvg::Path) {
    w.str(p.id());
    w.flag(p.is_visible());
    w.u32(matches!(p.paint_order(), usvg::PaintOrder::StrokeAndFill) as u32);
    w.flag(p.rendering_mode().use_shape_antialiasing());
    let bb = p.abs_stroke_bounding_box();
    w.flag(bb.width() > 0.0 && bb.height() > 0.0); // Node::abs_layer_bounding_box: to_non_zero_rect()
    w.rect(bb.x(), bb.y(), bb.width(), bb.height());
    w.flag(p.fill().is_some());
    if let Some(f) = p.fill() {
        write_paint(w, f.paint());
        w.f32(f.opacity().get());
        w.u32(matches!(f.rule(), usvg::FillRule::EvenOdd) as u32);
    }
    w.flag(p.stroke().is_some());
    if let Some(s) = p.stroke() {
        write_paint(w, s.paint());
        w.f32(s.opacity().get());
        w.f32(s.width().get());
        w.f32(s.miterlimit().get());
        w.u32(match s.linecap() { usvg::LineCap::Butt => 0, usvg::LineCap::Round => 1, usvg::LineCap::Square => 2 });
        w.u32(match s.linejoin() { usvg::LineJoin::Miter => 0, usvg::LineJoin::MiterClip => 1, usvg::LineJoin::Round => 2, usvg::LineJoin::Bevel => 3 });
        let dash = s.dasharray().unwrap_or(&[]);
        w.u32(dash.len() as u32);
        for d in dash { w.f32(*d) }
        w.f32(s.dashoffset());
    }
    let data = p.data();
    let verbs: Vec<u8> = data.verbs().iter().map(|v| match v {
        tiny_skia::PathVerb::Move => 0, tiny_skia::PathVerb::Line => 1, tiny_skia::PathVerb::Quad => 2,
        tiny_skia::PathVerb::Cubic => 3, tiny_skia::PathVerb::Close => 4,
    }).collect();
    w.u32(verbs.len() as u32);
    w.raw(&verbs);
    w.u32(data.points().len() as u32);
    for pt in data.points() { w.f32(pt.x); w.f32(pt.y) }
}

fn write_image(w: &mut W, im: &usvg::Image) {
    w.str(im.id());
    w.flag(im.is_visible());
    w.u32(match im.rendering_mode() { // image.rs:180-187
        usvg::ImageRendering::OptimizeQuality | usvg::ImageRendering::HighQuality => 2,
        usvg::ImageRendering::Smooth => 1,
        _ => 0,
    });
    let bb = im.abs_bounding_box();
    w.flag(bb.width() > 0.0 && bb.height() > 0.0);
    w.rect(bb.x(), bb.y(), bb.width(), bb.height());
    match im.kind() {
        usvg::ImageKind::SVG(sub) => { w.u32(0); write_tree(w, sub) }
        other => match decode_raster(other) {
            // decoded exactly as image.rs:62-170 (premultiplied RGBA8); an undecodable image draws nothing: an invisible 1x1
            Some(pm) => { w.u32(1); w.u32(pm.width()); w.u32(pm.height()); w.raw(pm.data()) }
            None => { w.u32(1); w.u32(1); w.u32(1); w.raw(&[0, 0, 0, 0]) }
        },
    }
}

fn decode_raster(kind: &usvg::ImageKind) -> Option<tiny_skia::Pixmap> {
    match kind {
        usvg::ImageKind::PNG(data) => tiny_skia::Pixmap::decode_png(data).ok(),
        // JPEG / GIF / WebP: the decoders of crates/resvg/src/image.rs:79-158 (zune-jpeg, gif, image-webp) move here unchanged.
        _ => None,
    }
}

fn write_input(w: &mut W, i: &usvg::filter::Input) {
    match i {
        usvg::filter::Input::SourceGraphic => w.u32(0),
        usvg::filter::Input::SourceAlpha => w.u32(1),
        usvg::filter::Input::Reference(name) => { w.u32(2); w.str(name) }
    }
}

fn write_transfer(w: &mut W, f: &usvg::filter::TransferFunction) {
    use usvg::filter::TransferFunction as T;
    let (ty, vals, p): (u32, &[f32], [f32; 5]) = match f {
        T::Identity => (0, &[], [1.0, 0.0, 1.0, 1.0, 0.0]),
        T::Table(v) => (1, v, [1.0, 0.0, 1.0, 1.0, 0.0]),
        T::Discrete(v) => (2, v, [1.0, 0.0, 1.0, 1.0, 0.0]),
        T::Linear { slope, intercept } => (3, &[], [*slope, *intercept, 1.0, 1.0, 0.0]),
        T::Gamma { amplitude, exponent, offset } => (4, &[], [1.0, 0.0, *amplitude, *exponent, *offset]),
    };
    w.u32(ty);
    w.u32(vals.len() as u32);
    for v in vals { w.f32(*v) }
    for v in p { w.f32(v) }
}

fn write_light(w: &mut W, l: usvg::filter::LightSource) {
    use usvg::filter::LightSource as L;
    let (kind, v, cone): (u32, [f32; 9], Option<f32>) = match l {
        L::DistantLight(d) => (0, [d.azimuth, d.elevation, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0], None),
        L::PointLight(p) => (1, [0.0, 0.0, p.x, p.y, p.z, 0.0, 0.0, 0.0, 1.0], None),
        L::SpotLight(s) => (2, [0.0, 0.0, s.x, s.y, s.z, s.points_at_x, s.points_at_y, s.points_at_z, s.specular_exponent.get()], s.limiting_cone_angle),
    };
    w.u32(kind);
    for x in v { w.f32(x) }
    w.flag(cone.is_some());
    w.f32(cone.unwrap_or(0.0));
}

fn write_primitive(w: &mut W, p: &usvg::filter::Primitive) {
    use usvg::filter::Kind as K;
    w.nz(p.rect());
    w.u32(matches!(p.color_interpolation(), usvg::filter::ColorInterpolation::LinearRGB) as u32);
    w.str(p.result());
    match p.kind() {
        K::Blend(fe) => { w.u32(0); w.u32(fe.mode() as u32); write_input(w, fe.input1()); write_input(w, fe.input2()) }
        K::DropShadow(fe) => {
            w.u32(1);
            write_input(w, fe.input());
            for v in [fe.dx(), fe.dy(), fe.std_dev_x().get(), fe.std_dev_y().get()] { w.f32(v) }
            w.rgb(fe.color());
            w.f32(fe.opacity().get());
        }
        K::Flood(fe) => { w.u32(2); w.rgb(fe.color()); w.f32(fe.opacity().get()) }
        K::GaussianBlur(fe) => { w.u32(3); write_input(w, fe.input()); w.f32(fe.std_dev_x().get()); w.f32(fe.std_dev_y().get()) }
        K::Offset(fe) => { w.u32(4); write_input(w, fe.input()); w.f32(fe.dx()); w.f32(fe.dy()) }
        K::Composite(fe) => {
            use usvg::filter::CompositeOperator as Op;
            w.u32(5);
            let (op, k) = match fe.operator() {
                Op::Over => (0, [0.0; 4]), Op::In => (1, [0.0; 4]), Op::Out => (2, [0.0; 4]), Op::Atop => (3, [0.0; 4]), Op::Xor => (4, [0.0; 4]),
                Op::Arithmetic { k1, k2, k3, k4 } => (5, [k1, k2, k3, k4]),
            };
            w.u32(op);
            for v in k { w.f32(v) }
            write_input(w, fe.input1());
            write_input(w, fe.input2());
        }
        K::Merge(fe) => { w.u32(6); w.u32(fe.inputs().len() as u32); for i in fe.inputs() { write_input(w, i) } }
        K::Tile(fe) => { w.u32(7); write_input(w, fe.input()) }
        K::Image(fe) => { w.u32(8); write_group(w, fe.root()) }
        K::ComponentTransfer(fe) => {
            w.u32(9);
            write_input(w, fe.input());
            for f in [fe.func_r(), fe.func_g(), fe.func_b(), fe.func_a()] { write_transfer(w, f) }
        }
        K::ColorMatrix(fe) => {
            use usvg::filter::ColorMatrixKind as M;
            w.u32(10);
            write_input(w, fe.input());
            match fe.kind() {
                M::Matrix(v) => { w.u32(0); w.u32(v.len() as u32); for x in v { w.f32(*x) } }
                M::Saturate(v) => { w.u32(1); w.u32(1); w.f32(v.get()) }
                M::HueRotate(v) => { w.u32(2); w.u32(1); w.f32(*v) }
                M::LuminanceToAlpha => { w.u32(3); w.u32(0) }
            }
        }
        K::ConvolveMatrix(fe) => {
            w.u32(11);
            write_input(w, fe.input());
            let m = fe.matrix();
            for v in [m.columns(), m.rows(), m.target_x(), m.target_y()] { w.u32(v) }
            w.f32(fe.divisor().get());
            w.f32(fe.bias());
            w.u32(match fe.edge_mode() { usvg::filter::EdgeMode::None => 0, usvg::filter::EdgeMode::Duplicate => 1, usvg::filter::EdgeMode::Wrap => 2 });
            w.flag(fe.preserve_alpha());
            w.u32(m.data().len() as u32);
            for v in m.data() { w.f32(*v) }
        }
        K::Morphology(fe) => {
            w.u32(12);
            write_input(w, fe.input());
            w.u32(matches!(fe.operator(), usvg::filter::MorphologyOperator::Dilate) as u32);
            w.f32(fe.radius_x().get());
            w.f32(fe.radius_y().get());
        }
        K::DisplacementMap(fe) => {
            w.u32(13);
            write_input(w, fe.input1());
            write_input(w, fe.input2());
            w.f32(fe.scale());
            w.u32(fe.x_channel_selector() as u32); // R, G, B, A
            w.u32(fe.y_channel_selector() as u32);
        }
        K::Turbulence(fe) => {
            w.u32(14);
            w.f32(fe.base_frequency_x().get());
            w.f32(fe.base_frequency_y().get());
            w.u32(fe.num_octaves());
            w.i32(fe.seed());
            w.flag(fe.stitch_tiles());
            w.flag(matches!(fe.kind(), usvg::filter::TurbulenceKind::FractalNoise));
        }
        K::DiffuseLighting(fe) => {
            w.u32(15);
            write_input(w, fe.input());
            for v in [fe.surface_scale(), fe.diffuse_constant(), 1.0] { w.f32(v) }
            w.rgb(fe.lighting_color());
            write_light(w, fe.light_source());
        }
        K::SpecularLighting(fe) => {
            w.u32(16);
            write_input(w, fe.input());
            for v in [fe.surface_scale(), fe.specular_constant(), fe.specular_exponent()] { w.f32(v) }
            w.rgb(fe.lighting_color());
            write_light(w, fe.light_source());
        }
    }
}
