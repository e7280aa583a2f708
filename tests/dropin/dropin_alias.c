/* tests/dropin/dropin_alias.c -- TEST INFRASTRUCTURE ONLY.
 *
 * The eight strong symbols that take the place of the reference's own (weakened with objcopy in the object
 * compiled from the unmodified nuts333.c): every call site in nuts333.c reaches its callee through the symbol
 * (-fPIC: R_X86_64_PLT32 relocations, no inlining across an interposable symbol), so after the link all of
 * them land here, in the bodies of shim/nuts333_shim.c.  Pointers only: this unit cannot see nuts333.h
 * (the header defines the talker's globals). */
void nutsb_shim_write_sock(int, char *);
void nutsb_shim_write_user(void *, char *);
void nutsb_shim_write_level(int, int, char *, void *);
void nutsb_shim_write_room(void *, char *);
void nutsb_shim_write_room_except(void *, char *, void *);
int  nutsb_shim_contains_swearing(char *);
int  nutsb_shim_site_banned(char *);
int  nutsb_shim_user_banned(char *);

void write_sock(int sock, char *str) { nutsb_shim_write_sock(sock, str); }
void write_user(void *user, char *str) { nutsb_shim_write_user(user, str); }
void write_level(int level, int above, char *str, void *user) { nutsb_shim_write_level(level, above, str, user); }
void write_room(void *rm, char *str) { nutsb_shim_write_room(rm, str); }
void write_room_except(void *rm, char *str, void *user) { nutsb_shim_write_room_except(rm, str, user); }
int  contains_swearing(char *str) { return nutsb_shim_contains_swearing(str); }
int  site_banned(char *site) { return nutsb_shim_site_banned(site); }
int  user_banned(char *name) { return nutsb_shim_user_banned(name); }
