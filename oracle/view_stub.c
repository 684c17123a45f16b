/* oracle/view_stub.c -- TEST INFRASTRUCTURE. The reference's -DUNIX_INT64 configuration calls
 * sdpDataMatView() (lorads_sdp_conic.c:663) which is defined nowhere in the tree; this empty
 * definition lets the 64-bit oracle binary link. It is only reached by a debug cone printer. */
void sdpDataMatView(void *p) { (void)p; }
