//! babyjubjub-rs API over the B200 batch engine (libbjj_cuda).
//!
//! UNTESTED SOURCE: the build image has no Rust toolchain.  The same C ABI is exercised by the Python
//! (`babyjubjub-rs_b200/`) and C++ (`host/bjj.hpp`) mirrors, which the parity tests drive.
//!
//! The public names follow arnaucube/babyjubjub-rs 0.0.11: `Point`, `PointProjective`, `Signature`,
//! `PrivateKey`, `decompress_point`, `decompress_signature`, `verify`, plus the batch entry points
//! `mul_scalar_batch`, `public_batch`, `decompress_batch`, `verify_batch`.  `Fr` is carried as its
//! canonical 32-byte little-endian value (what ff_ce's `into_repr()` yields).  There is no CPU
//! fallback: without a CUDA device `Engine::new` returns an error.
pub mod ffi;

use num_bigint::{BigInt, Sign};
use num_traits::Zero;
use std::ffi::CStr;
use std::sync::OnceLock;

pub type Fr = [u8; 32];

pub fn q() -> BigInt {
    BigInt::parse_bytes(b"21888242871839275222246405745257275088548364400416034343698204186575808495617", 10).unwrap()
}

pub fn fr_from_str(s: &str) -> Option<Fr> {
    let v = BigInt::parse_bytes(s.as_bytes(), 10)?;
    if v.sign() == Sign::Minus {
        return None;
    }
    Some(bigint_le32(&(v % q())))
}

/// 32 little-endian bytes of |v| mod 2^256 (scalars are never reduced by the engine: reference semantics)
pub fn bigint_le32(v: &BigInt) -> [u8; 32] {
    let (_, bytes) = v.to_bytes_le();
    let mut out = [0u8; 32];
    let n = bytes.len().min(32);
    out[..n].copy_from_slice(&bytes[..n]);
    out
}

fn msg_le32(msg: &BigInt) -> [u8; 32] {
    // `verify` returns false for msg > Q; any 32-byte value > Q carries that through the ABI
    if msg > &q() {
        [0xFF; 32]
    } else {
        bigint_le32(msg)
    }
}

pub struct Engine {
    ctx: *mut ffi::bjj_ctx,
}
unsafe impl Send for Engine {}
unsafe impl Sync for Engine {}

impl Engine {
    pub fn new(device: i32) -> Result<Engine, String> {
        let mut ctx = std::ptr::null_mut();
        let rc = unsafe { ffi::bjj_init(device, &mut ctx) };
        if rc != ffi::BJJ_OK {
            return Err(format!("bjj_init failed: {} (there is no CPU fallback)", err_str(rc)));
        }
        Ok(Engine { ctx })
    }
    pub fn raw(&self) -> *mut ffi::bjj_ctx {
        self.ctx
    }
}
impl Drop for Engine {
    fn drop(&mut self) {
        unsafe { ffi::bjj_destroy(self.ctx) }
    }
}

fn err_str(rc: i32) -> String {
    unsafe { CStr::from_ptr(ffi::bjj_error_string(rc)).to_string_lossy().into_owned() }
}
fn status_str(st: u8) -> String {
    unsafe { CStr::from_ptr(ffi::bjj_status_string(st as i32)).to_string_lossy().into_owned() }
}
fn check(rc: i32, what: &str) {
    assert!(rc == ffi::BJJ_OK, "{} failed: {}", what, err_str(rc));
}

static ENGINE: OnceLock<Engine> = OnceLock::new();
pub fn engine() -> &'static Engine {
    ENGINE.get_or_init(|| Engine::new(0).expect("CUDA device required"))
}

fn col<F: Fn(usize) -> [u8; 32]>(n: usize, f: F) -> Vec<u8> {
    let mut v = Vec::with_capacity(32 * n);
    for i in 0..n {
        v.extend_from_slice(&f(i));
    }
    v
}
fn row(buf: &[u8], i: usize) -> [u8; 32] {
    let mut r = [0u8; 32];
    r.copy_from_slice(&buf[32 * i..32 * i + 32]);
    r
}

#[derive(Clone, Debug, PartialEq)]
pub struct PointProjective {
    pub x: Fr,
    pub y: Fr,
    pub z: Fr,
}

#[derive(Clone, Debug, PartialEq)]
pub struct Point {
    pub x: Fr,
    pub y: Fr,
}

impl PointProjective {
    pub fn affine(&self) -> Point {
        let (mut rx, mut ry) = ([0u8; 32], [0u8; 32]);
        check(unsafe { ffi::bjj_affine_batch(engine().raw(), 1, self.x.as_ptr(), self.y.as_ptr(), self.z.as_ptr(),
                                             rx.as_mut_ptr(), ry.as_mut_ptr()) }, "bjj_affine_batch");
        Point { x: rx, y: ry }
    }
    pub fn add(&self, q: &PointProjective) -> PointProjective {
        let (mut rx, mut ry, mut rz) = ([0u8; 32], [0u8; 32], [0u8; 32]);
        check(unsafe { ffi::bjj_add_batch(engine().raw(), 1, self.x.as_ptr(), self.y.as_ptr(), self.z.as_ptr(),
                                          q.x.as_ptr(), q.y.as_ptr(), q.z.as_ptr(), rx.as_mut_ptr(), ry.as_mut_ptr(),
                                          rz.as_mut_ptr()) }, "bjj_add_batch");
        PointProjective { x: rx, y: ry, z: rz }
    }
}

impl Point {
    pub fn projective(&self) -> PointProjective {
        let mut one = [0u8; 32];
        one[0] = 1;
        PointProjective { x: self.x, y: self.y, z: one }
    }
    pub fn mul_scalar(&self, n: &BigInt) -> Point {
        mul_scalar_batch(std::slice::from_ref(self), std::slice::from_ref(n)).pop().unwrap()
    }
    pub fn compress(&self) -> [u8; 32] {
        let mut out = [0u8; 32];
        check(unsafe { ffi::bjj_compress_batch(engine().raw(), 1, self.x.as_ptr(), self.y.as_ptr(), out.as_mut_ptr()) },
              "bjj_compress_batch");
        out
    }
    pub fn equals(&self, p: Point) -> bool {
        self.x == p.x && self.y == p.y
    }
}

pub fn decompress_point(bb: [u8; 32]) -> Result<Point, String> {
    decompress_batch(&[bb]).pop().unwrap()
}

#[derive(Debug, Clone)]
pub struct Signature {
    pub r_b8: Point,
    pub s: BigInt,
}
impl Signature {
    pub fn compress(&self) -> [u8; 64] {
        let mut r = [0u8; 64];
        r[..32].copy_from_slice(&self.r_b8.compress());
        r[32..].copy_from_slice(&bigint_le32(&self.s));
        r
    }
}
pub fn decompress_signature(b: &[u8; 64]) -> Result<Signature, String> {
    let mut rb = [0u8; 32];
    rb.copy_from_slice(&b[..32]);
    let r_b8 = decompress_point(rb)?;
    Ok(Signature { r_b8, s: BigInt::from_bytes_le(Sign::Plus, &b[32..]) })
}

pub struct PrivateKey {
    pub key: [u8; 32],
}
impl PrivateKey {
    pub fn import(b: Vec<u8>) -> Result<PrivateKey, String> {
        if b.len() != 32 {
            return Err(String::from("imported key can not be bigger than 32 bytes"));
        }
        let mut key = [0u8; 32];
        key.copy_from_slice(&b);
        Ok(PrivateKey { key })
    }
    pub fn scalar_key(&self) -> BigInt {
        let mut out = [0u8; 32];
        check(unsafe { ffi::bjj_scalar_key_batch(engine().raw(), 1, self.key.as_ptr(), out.as_mut_ptr()) },
              "bjj_scalar_key_batch");
        BigInt::from_bytes_le(Sign::Plus, &out)
    }
    pub fn public(&self) -> Point {
        public_batch(std::slice::from_ref(self)).pop().unwrap()
    }
    pub fn sign(&self, msg: BigInt) -> Result<Signature, String> {
        if msg > q() {
            return Err("msg outside the Finite Field".to_string());
        }
        let m = bigint_le32(&msg);
        let (mut rx, mut ry, mut s, mut st) = ([0u8; 32], [0u8; 32], [0u8; 32], 0u8);
        check(unsafe { ffi::bjj_sign_batch(engine().raw(), 1, self.key.as_ptr(), m.as_ptr(), rx.as_mut_ptr(),
                                           ry.as_mut_ptr(), s.as_mut_ptr(), &mut st) }, "bjj_sign_batch");
        if st != 0 {
            return Err(status_str(st));
        }
        Ok(Signature { r_b8: Point { x: rx, y: ry }, s: BigInt::from_bytes_le(Sign::Plus, &s) })
    }
}

pub fn verify(pk: Point, sig: Signature, msg: BigInt) -> bool {
    verify_batch(&[pk], &[sig], &[msg])[0]
}

// ---- batch entry points ------------------------------------------------------------------------------

/// `Point::mul_scalar` over a batch.  Scalars wider than 256 bits must be reduced by the caller
/// (mod ORDER, and only for points that are on the curve).
pub fn mul_scalar_batch(points: &[Point], scalars: &[BigInt]) -> Vec<Point> {
    let n = points.len();
    assert_eq!(n, scalars.len());
    let (px, py) = (col(n, |i| points[i].x), col(n, |i| points[i].y));
    let k = col(n, |i| { assert!(scalars[i].bits() <= 256); bigint_le32(&scalars[i]) });
    let (mut rx, mut ry) = (vec![0u8; 32 * n], vec![0u8; 32 * n]);
    check(unsafe { ffi::bjj_mul_scalar_batch(engine().raw(), n, px.as_ptr(), py.as_ptr(), k.as_ptr(), rx.as_mut_ptr(),
                                             ry.as_mut_ptr()) }, "bjj_mul_scalar_batch");
    (0..n).map(|i| Point { x: row(&rx, i), y: row(&ry, i) }).collect()
}

pub fn public_batch(keys: &[PrivateKey]) -> Vec<Point> {
    let n = keys.len();
    let k = col(n, |i| keys[i].key);
    let (mut rx, mut ry) = (vec![0u8; 32 * n], vec![0u8; 32 * n]);
    check(unsafe { ffi::bjj_public_batch(engine().raw(), n, k.as_ptr(), rx.as_mut_ptr(), ry.as_mut_ptr()) },
          "bjj_public_batch");
    (0..n).map(|i| Point { x: row(&rx, i), y: row(&ry, i) }).collect()
}

pub fn decompress_batch(blobs: &[[u8; 32]]) -> Vec<Result<Point, String>> {
    let n = blobs.len();
    let inp = col(n, |i| blobs[i]);
    let (mut rx, mut ry, mut st) = (vec![0u8; 32 * n], vec![0u8; 32 * n], vec![0u8; n]);
    check(unsafe { ffi::bjj_decompress_batch(engine().raw(), n, inp.as_ptr(), rx.as_mut_ptr(), ry.as_mut_ptr(),
                                             st.as_mut_ptr()) }, "bjj_decompress_batch");
    (0..n).map(|i| if st[i] == 0 { Ok(Point { x: row(&rx, i), y: row(&ry, i) }) } else { Err(status_str(st[i])) }).collect()
}

pub fn verify_batch(pks: &[Point], sigs: &[Signature], msgs: &[BigInt]) -> Vec<bool> {
    let n = pks.len();
    assert!(n == sigs.len() && n == msgs.len());
    let (r8x, r8y) = (col(n, |i| sigs[i].r_b8.x), col(n, |i| sigs[i].r_b8.y));
    let s = col(n, |i| { assert!(!sigs[i].s.is_zero() || true); bigint_le32(&sigs[i].s) });
    let (ax, ay) = (col(n, |i| pks[i].x), col(n, |i| pks[i].y));
    let m = col(n, |i| msg_le32(&msgs[i]));
    let mut ok = vec![0u8; n];
    check(unsafe { ffi::bjj_verify_batch(engine().raw(), n, r8x.as_ptr(), r8y.as_ptr(), s.as_ptr(), ax.as_ptr(), ay.as_ptr(),
                                         m.as_ptr(), ok.as_mut_ptr()) }, "bjj_verify_batch");
    ok.into_iter().map(|b| b == 1).collect()
}
