/* Java side of the direct-ByteBuffer and persistent-key natives of liboctozk's VariableBaseMSM shim
 * (octopuszk_b200/csrc/jni/jni_shim.cc).  UNVERIFIED: the build image has no JDK, this file has never been compiled; it is
 * the source a maintainer would add next to the reference's algebra/msm/VariableBaseMSM.java (see INTEGRATION.md).
 *
 * What it replaces: the per-element marshalling of VariableBaseMSM.serialMSM (VariableBaseMSM.java:217-237: one
 * BigInteger.toByteArray + ByteArrayOutputStream.write per coordinate, then a byte[] copy) by one pass into a direct
 * ByteBuffer of 32-byte little-endian elements, and the re-marshalling of the proving-key vectors on every proof
 * (SerialProver.java:70-106) by handles to device-resident bases.
 *
 * The native methods live in class algebra.msm.VariableBaseMSM (that is the JNI name the shim exports), so the three
 * declarations below must be added to that class:
 *
 *     public static native int  variableBaseMSMDirect(ByteBuffer bases1, ByteBuffer bases2, ByteBuffer scalars,
 *                                                      int batch_size, int type, int taskID, ByteBuffer out);
 *     public static native long uploadBasesDirect(ByteBuffer bases, int count, int type, int taskID);
 *     public static native void freeBases(long handle, int taskID);
 *     public static native int  variableBaseMSMKeyedDirect(long keyG1, long keyG2, ByteBuffer scalars, int first,
 *                                                           int batch_size, int type, int taskID, ByteBuffer out);
 */
package algebra.msm;

import java.math.BigInteger;
import java.nio.ByteBuffer;
import java.nio.ByteOrder;
import java.util.List;

public final class OctozkDirect {
    private OctozkDirect() {}

    /** value (0 <= value < 2^256) as 32 little-endian bytes at the buffer's position. */
    static void putLE32(final ByteBuffer buf, final BigInteger value) {
        final byte[] be = value.toByteArray();                 // big-endian, possibly with a leading sign byte
        final int n = Math.min(be.length, 32);
        for (int j = 0; j < n; j++) buf.put(be[be.length - 1 - j]);
        for (int j = n; j < 32; j++) buf.put((byte) 0);
    }

    /** 32 little-endian bytes at the buffer's position as a non-negative BigInteger. */
    static BigInteger getLE32(final ByteBuffer buf) {
        final byte[] be = new byte[33];                        // be[0] = 0: positive
        for (int j = 0; j < 32; j++) be[32 - j] = buf.get();
        return new BigInteger(be);
    }

    /** n x 32 bytes: the scalars as the natives expect them. */
    public static ByteBuffer packScalars(final List<BigInteger> scalars) {
        final ByteBuffer buf = ByteBuffer.allocateDirect(32 * scalars.size()).order(ByteOrder.LITTLE_ENDIAN);
        for (final BigInteger s : scalars) putLE32(buf, s);
        buf.rewind();
        return buf;
    }

    /** n x 96 bytes (X|Y|Z) from the BN254G1ToBigInteger() triples of the bases (VariableBaseMSM.java:224-227). */
    public static ByteBuffer packG1(final List<? extends List<BigInteger>> xyz) {
        final ByteBuffer buf = ByteBuffer.allocateDirect(96 * xyz.size()).order(ByteOrder.LITTLE_ENDIAN);
        for (final List<BigInteger> p : xyz) {
            putLE32(buf, p.get(0));
            putLE32(buf, p.get(1));
            putLE32(buf, p.get(2));
        }
        buf.rewind();
        return buf;
    }

    /** n x 192 bytes (X.c0|X.c1|Y.c0|Y.c1|Z.c0|Z.c1) from BN254G2ToBigInteger() (VariableBaseMSM.java:279-285). */
    public static ByteBuffer packG2(final List<? extends List<BigInteger>> coords) {
        final ByteBuffer buf = ByteBuffer.allocateDirect(192 * coords.size()).order(ByteOrder.LITTLE_ENDIAN);
        for (final List<BigInteger> p : coords)
            for (int k = 0; k < 6; k++) putLE32(buf, p.get(k));
        buf.rewind();
        return buf;
    }

    /** sum_i scalars[i] * bases[i] in G1; returns {X, Y, Z} (Jacobian, fully reduced; (0,1,0) is infinity). */
    public static BigInteger[] msmG1(final ByteBuffer scalars, final ByteBuffer bases, final int n, final int taskID) {
        final ByteBuffer out = ByteBuffer.allocateDirect(96).order(ByteOrder.LITTLE_ENDIAN);
        VariableBaseMSM.variableBaseMSMDirect(bases, null, scalars, n, 1, taskID, out);    // throws RuntimeException on error
        return new BigInteger[] {getLE32(out), getLE32(out), getLE32(out)};
    }

    /** A proving-key vector kept on the device: upload once (after setup or after deserialising the key), use per proof. */
    public static final class ResidentBases implements AutoCloseable {
        private long handle;
        private final int type, taskID, count;

        public ResidentBases(final ByteBuffer bases, final int count, final int type, final int taskID) {
            this.handle = VariableBaseMSM.uploadBasesDirect(bases, count, type, taskID);
            this.type = type;
            this.taskID = taskID;
            this.count = count;
        }

        /** sum_{i < n} scalars[i] * key[first + i]: G1 returns 3 coordinates, G2 returns 6. */
        public BigInteger[] msm(final ByteBuffer scalars, final int first, final int n) {
            if (first < 0 || n < 0 || first + n > count) throw new IllegalArgumentException("range exceeds the uploaded bases");
            final int words = type == 1 ? 3 : 6;
            final ByteBuffer out = ByteBuffer.allocateDirect(32 * words).order(ByteOrder.LITTLE_ENDIAN);
            VariableBaseMSM.variableBaseMSMKeyedDirect(type == 1 ? handle : 0L, type == 2 ? handle : 0L, scalars, first, n, type, taskID, out);
            final BigInteger[] r = new BigInteger[words];
            for (int k = 0; k < words; k++) r[k] = getLE32(out);
            return r;
        }

        @Override
        public void close() {
            if (handle != 0L) VariableBaseMSM.freeBases(handle, taskID);
            handle = 0L;
        }
    }
}
