//! `extern "C"` declarations of include/resvg_b200.h — the whole-tree entry points and what they need.
//! (The per-call seam — rb_fill_path, rb_draw_layer, the filter primitives ... — is declared in the header as well; the
//! shim does not need it: the traversal runs inside the library.)
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct rb_ctx { _p: [u8; 0] }
#[repr(C)]
pub struct rb_layer { _p: [u8; 0] }
#[repr(C)]
pub struct rb_tree { _p: [u8; 0] }

pub const RB_OK: c_int = 0;
pub const RB_ERR_INVALID: c_int = 1;

extern "C" {
    pub fn rb_ctx_create(device: c_int, out: *mut *mut rb_ctx) -> c_int;
    pub fn rb_ctx_destroy(ctx: *mut rb_ctx);
    pub fn rb_ctx_synchronize(ctx: *mut rb_ctx) -> c_int;
    pub fn rb_last_error(ctx: *mut rb_ctx) -> *const c_char;

    pub fn rb_layer_create(ctx: *mut rb_ctx, width: u32, height: u32, out: *mut *mut rb_layer) -> c_int;
    pub fn rb_layer_destroy(layer: *mut rb_layer);
    pub fn rb_layer_upload(layer: *mut rb_layer, host_rgba: *const u8) -> c_int;
    pub fn rb_layer_download(layer: *mut rb_layer, host_rgba: *mut u8) -> c_int;

    pub fn rb_tree_parse(stream: *const c_void, len: usize, out: *mut *mut rb_tree) -> c_int;
    pub fn rb_tree_destroy(tree: *mut rb_tree);
    pub fn rb_tree_node_bbox(tree: *const rb_tree, id: *const c_char, out_xywh: *mut f32) -> c_int;
    pub fn rb_render(ctx: *mut rb_ctx, tree: *const rb_tree, ts: *const f32, target: *mut rb_layer) -> c_int;
    pub fn rb_render_node(ctx: *mut rb_ctx, tree: *const rb_tree, id: *const c_char, ts: *const f32, target: *mut rb_layer) -> c_int;
    pub fn rb_render_to_host(ctx: *mut rb_ctx, tree: *const rb_tree, ts: *const f32, width: u32, height: u32, pixmap: *mut u8) -> c_int;
}
