// Links libronk_b200.so (built by `make -C ronkathon_b200/csrc`); RONK_B200_LIB_DIR points at it.
fn main() {
  let dir = std::env::var("RONK_B200_LIB_DIR").unwrap_or_else(|_| "../../../ronkathon_b200".into());
  println!("cargo:rustc-link-search=native={dir}");
  println!("cargo:rustc-link-lib=dylib=ronk_b200");
}
