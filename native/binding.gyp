{
  "targets": [
    {
      "target_name": "pzk",
      "sources": ["pzk_napi.cc"],
      "include_dirs": ["../include"],
      "libraries": ["-L<(module_root_dir)/../passport-zk-circuits_b200/lib", "-lpzk",
                    "-Wl,-rpath,<(module_root_dir)/../passport-zk-circuits_b200/lib"],
      "cflags_cc": ["-std=c++17", "-O2"],
      "defines": ["NAPI_VERSION=8"]
    }
  ]
}
